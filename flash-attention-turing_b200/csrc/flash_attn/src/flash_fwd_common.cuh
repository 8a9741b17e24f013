// flash_fwd_common.cuh — parameter block, tile constants and the persistent work list of the forward kernel
// (flash_fwd_p4_sm100.cu) and its host launcher (flash_fwd_sm100.cu).
#pragma once
#include <stdlib.h>

#include "fa_common.h"
#include "sm100_ptx.cuh"

namespace fa100 {

struct FwdParams {
    void* o;
    float* lse;
    const int* cu_q;
    const int* cu_k;
    int b, sq, sk, h, h_k, hratio;
    int is_causal;
    int exact;         // 1: every key step computes the exact row max first (no speculative steps); FA_B200_FWD_EXACT
    float scale;       // 1/sqrt(d)
    float scale_log2;  // log2(e)/sqrt(d)
    float inv_scale_log2;
    long long* trace;  // FA_TRACE builds only: clock64() stamps of CTA 0, [8 roles][64 steps][8 events]
};

#ifdef FA_TRACE
#define FA_TRACE_EVENT(role, j, ev)                                                              \
    do {                                                                                         \
        if (p.trace && blockIdx.x == 0 && (role) < 8 && (j) < 64)                                 \
            p.trace[((role) * 64 + (j)) * 8 + (ev)] = clock64();                                 \
    } while (0)
#else
#define FA_TRACE_EVENT(role, j, ev) do {} while (0)
#endif

constexpr int kBlockM = 128;  // query rows per tile  (= TMEM lanes = UMMA M)
constexpr int kBlockN = 128;  // key rows per tile    (= UMMA N of S, K extent of PV)

constexpr uint32_t kTmemS0 = 0, kTmemO0 = 256;      // tile t: S at kTmemS0 + 128 t, O at kTmemO0 + 128 t
constexpr float kRescaleThreshold = 8.0f;


// ---- persistent-kernel work list ----------------------------------------------------------------------
// division by a runtime constant as multiply-high + shift (valid for numerators < 2^31): the work list is decoded by
// every role at every item, and three 32-bit divisions cost ~250 cycles of dependent integer code each time
struct FastDiv {
    uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d; f.mul = 0; f.shr = 0;
    if (d > 1) {
        uint32_t lg = 0;
        while ((1ull << lg) < d) ++lg;                       // ceil(log2 d)
        const uint32_t pw = 31 + lg;
        f.mul = (uint32_t)(((1ull << pw) + d - 1) / d);
        f.shr = pw - 32;
    }
    return f;
}
FA_DEVICE int fast_div(int n, const FastDiv& f) { return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr); }

struct TileSched {
    int num_mblk;      // 256-row query blocks per (batch, head)
    int levels;        // schedule slots per (batch, head): num_mblk, or ceil(num_mblk / 2) when row blocks are paired
    int paired;        // causal: a schedule slot holds TWO row blocks of one head, level L and num_mblk - 1 - L
    int bh;            // batch * heads
    int group;         // (batch, head) pairs per L2 group
    int total;         // schedule slots: levels * bh
    FastDiv per_group; // group * levels
    FastDiv full;      // heads in a full group (= group)
    FastDiv last;      // heads in the last, possibly smaller group
    FastDiv heads;     // h
};

struct WorkItem {
    int mblk, bidh, bidb;
};

// Static work list.  Schedule slots are dealt round-robin to the CTAs (slot = blockIdx + k * gridDim); inside an L2 group
// of heads the slots go level by level with the heads interleaved, so the CTAs that run concurrently work on the same few
// heads.  Causal row blocks cost 2 m + 2 key steps (m = row-block index): a slot pairs the heavy block num_mblk - 1 - L
// with the light block L of the same head — every slot costs the same 2 num_mblk + 2 steps, the static deal is balanced
// without atomics (round 1 dealt single row blocks heaviest-first: the per-CTA sums differed by ~10 % at s = 16384 and a
// causal forward took 0.63 of the non-causal time instead of ~0.52).
// An item code n is the slot, or 2 * slot + half for paired schedules (half 0 = the heavy block, 1 = its light partner).
FA_DEVICE WorkItem decode_item(const TileSched& ts, int n, int h, bool causal) {
    (void)causal;
    const int half = ts.paired ? (n & 1) : 0;
    const int slot = ts.paired ? (n >> 1) : n;
    const int g = fast_div(slot, ts.per_group);
    const int r = slot - g * (int)ts.per_group.d;
    const bool is_last = (g + 1) * ts.group > ts.bh;              // the last group may be smaller
    const int heads_here = is_last ? (int)ts.last.d : ts.group;
    const int level = is_last ? fast_div(r, ts.last) : fast_div(r, ts.full);
    const int head_local = r - level * heads_here;
    const int bhi = g * ts.group + head_local;
    WorkItem w;
    if (ts.paired) {
        const int heavy = ts.num_mblk - 1 - level;
        // odd num_mblk: the middle block has no partner -> the second half is a void item (row block beyond the sequence)
        w.mblk = half == 0 ? heavy : (level < heavy ? level : ts.num_mblk);
    } else {
        w.mblk = level;
    }
    w.bidb = fast_div(bhi, ts.heads);
    w.bidh = bhi - w.bidb * h;
    return w;
}

// geometry of one work item (identical in every role)
struct ItemGeom {
    int m0, q_row0, k_row0, sq_b, sk_b, tma_b, causal_off, nblk[2], n_blocks;
    bool skip;
};
FA_DEVICE ItemGeom item_geom(const FwdParams& p, const WorkItem& w) {
    ItemGeom g;
    g.m0 = w.mblk * (2 * kBlockM);
    if (p.cu_q != nullptr) {
        g.q_row0 = p.cu_q[w.bidb];
        g.sq_b = p.cu_q[w.bidb + 1] - g.q_row0;
        g.k_row0 = p.cu_k[w.bidb];
        g.sk_b = p.cu_k[w.bidb + 1] - g.k_row0;
        // max_seqlen_q sizes the LSE rows ([b, h, max_seqlen_q]): a sequence longer than the caller declared must not
        // write past its LSE row (the reference trusts the same input, flash_api.cpp:352-360); clamp instead
        g.sq_b = max(0, min(g.sq_b, p.sq));
        g.sk_b = max(0, min(g.sk_b, p.sk));
        g.tma_b = 0;
    } else {
        g.q_row0 = 0; g.k_row0 = 0; g.sq_b = p.sq; g.sk_b = p.sk; g.tma_b = w.bidb;
    }
    g.skip = g.m0 >= g.sq_b;
    g.causal_off = g.sk_b - g.sq_b;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int mt = g.m0 + t * kBlockM;
        int kv_end = (mt < g.sq_b) ? g.sk_b : 0;
        if (p.is_causal) kv_end = min(kv_end, max(0, mt + kBlockM + g.causal_off));
        g.nblk[t] = (kv_end + kBlockN - 1) / kBlockN;
    }
    g.n_blocks = max(g.nblk[0], g.nblk[1]);
    return g;
}

// static work list of the persistent kernels (host side): see the comment at decode_item
inline TileSched make_tile_sched(const fa_fwd_params* p) {
    TileSched ts;
    ts.num_mblk = (int)((p->seqlen_q + 2 * kBlockM - 1) / (2 * kBlockM));
    ts.bh = (int)(p->b * p->h);            // validate_fwd (fa_api.cu) bounds b * h * query blocks to 31 bits
    // (batch, head) pairs per group: their K and V (2 * sk * d * 2 bytes each pair of tensors... per KV head) should stay
    // L2-resident while the group is being worked on (default 16 MB of the 126 MB L2, see FA_B200_GROUP_MB below)
    const int64_t kv_bytes_per_head = 2 * p->seqlen_k * p->d * 2;
    static int64_t l2_budget_mb = -1;   // FA_B200_GROUP_MB: tuning knob for the L2 working set of one head group
    if (l2_budget_mb < 0) {
        const char* e = getenv("FA_B200_GROUP_MB");
        l2_budget_mb = e ? atoll(e) : 16;   // measured on B200: 12-24 MB best (C2 and C3), 48+ loses L2 locality
        if (l2_budget_mb < 1) l2_budget_mb = 1;
    }
    int64_t grp = kv_bytes_per_head > 0 ? (l2_budget_mb << 20) / kv_bytes_per_head : ts.bh;
    if (grp < 1) grp = 1;
    if (grp > ts.bh) grp = ts.bh;
    ts.group = (int)grp;
    ts.paired = p->is_causal ? 1 : 0;
    ts.levels = ts.paired ? (ts.num_mblk + 1) / 2 : ts.num_mblk;
    ts.total = ts.levels * ts.bh;
    ts.per_group = make_fastdiv((uint32_t)(ts.group * ts.levels));
    ts.full = make_fastdiv((uint32_t)ts.group);
    const int last_heads = ts.bh - (ts.bh - 1) / ts.group * ts.group;   // 1 .. group
    ts.last = make_fastdiv((uint32_t)last_heads);
    ts.heads = make_fastdiv((uint32_t)p->h);
    return ts;
}

// launcher of the forward kernel (flash_fwd_p4_sm100.cu); returns FA_OK / FA_ERR_*
template <int D, bool kBf16, int kEmu>
int launch_fwd_p4(const fa_fwd_params* p, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, FwdParams kp,
                  cudaStream_t stream);

}  // namespace fa100
