// flash_fwd_common.cuh — definitions shared by the forward kernels (flash_fwd_sm100.cu, flash_fwd_persist_sm100.cu).
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
    float scale;       // 1/sqrt(d)
    float scale_log2;  // log2(e)/sqrt(d)
    float inv_scale_log2;
    long long* trace;  // FA_TRACE builds only: clock64() stamps of CTA (0,0,0), [role][j][event]
};

#ifdef FA_TRACE
#define FA_TRACE_EVENT(role, j, ev)                                                              \
    do {                                                                                         \
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 64)        \
            p.trace[((role) * 64 + (j)) * 8 + (ev)] = clock64();                                 \
    } while (0)
#else
#define FA_TRACE_EVENT(role, j, ev) do {} while (0)
#endif

constexpr int kBlockM = 128;  // query rows per tile  (= TMEM lanes = UMMA M)
constexpr int kBlockN = 128;  // key rows per tile    (= UMMA N of S, K extent of PV)

template <int D> struct FwdSmem {
    static constexpr int kSlab = kBlockM * 128;      // 64-column slab of a 128-row tile: 16 KB
    static constexpr int kTile = kBlockM * D * 2;    // one 128 x D tile
    static constexpr int kKvStages = (D == 128) ? 4 : 6;
    static constexpr int kOffQ = 0;                  // 2 tiles
    static constexpr int kOffKV = 2 * kTile;         // ring of K/V tiles: K_j -> slot 2j, V_j -> slot 2j+1
    static constexpr int kOffBar = kOffKV + kKvStages * kTile;
    static constexpr int kBytes = kOffBar + 512 + 1024;
};
constexpr uint32_t kTmemS0 = 0, kTmemO0 = 256;      // tile t: S at kTmemS0 + 128 t, O at kTmemO0 + 128 t
constexpr float kRescaleThreshold = 8.0f;


// One online-softmax step for one query row (one thread) over a 128-key tile whose scores are already in registers
// (masked entries = -inf).  Writes P (16-bit) into the first 64 columns of the S tile in TMEM in two halves, arriving on
// bar_p0 / bar_p1 so that P V can start on the first half; updates the row's reference max m_ref and running sum l_run,
// rescaling O in TMEM only when the reference moves.
template <int D, bool kBf16, int kEmu>
FA_DEVICE void softmax_step(float (&s)[kBlockN], const bool first, const float c2, const float inv_c2, float& m_ref, float& l_run,
                            const uint32_t tS, const uint32_t tO, uint64_t* bar_p0, uint64_t* bar_p1) {
    // ---- online softmax with SPECULATIVE exponentials ----------------------------------------------
    // The reference max m_ref only moves when the tile max exceeds it by more than 2^8 (lazy rescale), so
    // for every tile but the first the exponentials can start immediately with the old reference while the
    // row max (3-input FMNMX on the ALU pipe) is computed underneath the MUFU-bound exp loop.  Only if the
    // vote afterwards says "rescale" (rare) is the first half recomputed with the new reference — the
    // values are bit-identical to computing the max first, the critical path is ~350 cycles shorter.
    if (first) {   // no reference yet: exact max first
        float mxa = fmaxf(s[0], s[1]), mxb = fmaxf(s[2], s[3]), mxc = fmaxf(s[4], s[5]), mxd = fmaxf(s[6], s[7]);
#pragma unroll
        for (int c = 8; c < kBlockN; c += 8) {
            mxa = fmaxf(mxa, fmaxf(s[c], s[c + 1]));
            mxb = fmaxf(mxb, fmaxf(s[c + 2], s[c + 3]));
            mxc = fmaxf(mxc, fmaxf(s[c + 4], s[c + 5]));
            mxd = fmaxf(mxd, fmaxf(s[c + 6], s[c + 7]));
        }
        m_ref = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd));
    }
    float neg = (m_ref == -INFINITY) ? 0.f : -m_ref * c2;
    const float2 c2v = make_float2(c2, c2);
    const float2 magic = make_float2(12582912.f, 12582912.f);       // 1.5 * 2^23
    // P = 2^(s*c2 + neg), two neighbouring columns at a time on the packed fp32x2 pipe (FFMA2 / FADD2).
    // kEmu of every 4 pairs use a Cody-Waite split + degree-3 minimax polynomial (|rel err| < 7.5e-5, far
    // below the 16-bit rounding of P) on the FMA pipe instead of MUFU.EX2.
    auto ex2_pair = [&](float a, float b, float negx, bool emulate) -> float2 {
        const float2 negv = make_float2(negx, negx);
        if (!emulate) {
            const float2 x = __ffma2_rn(make_float2(a, b), c2v, negv);
            return make_float2(fast_exp2(x.x), fast_exp2(x.y));
        }
        // clamp so that 2^x stays a normal float (above 126 the exponent add below would wrap around; the speculative
        // pass can see such arguments, the max vote then redoes the tile)
        const float s_floor = (-125.f - negx) * inv_c2, s_ceil = (126.f - negx) * inv_c2;
        const float2 x = __ffma2_rn(make_float2(fminf(fmaxf(a, s_floor), s_ceil), fminf(fmaxf(b, s_floor), s_ceil)), c2v, negv);
        const float2 tt = __fadd2_rn(x, magic);                   // low mantissa bits = rint(x)
        const float2 nnf = __ffma2_rn(tt, make_float2(-1.f, -1.f), magic);   // -rint(x), exact
        const float2 f = __fadd2_rn(x, nnf);                      // x - rint(x)  in [-0.5, 0.5]
        float2 pl = __ffma2_rn(make_float2(0.05517115816473961f, 0.05517115816473961f), f,
                               make_float2(0.2426101416349411f, 0.2426101416349411f));
        pl = __ffma2_rn(pl, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
        pl = __ffma2_rn(pl, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
        return make_float2(__uint_as_float(__float_as_uint(pl.x) + (__float_as_uint(tt.x) << 23)),
                           __uint_as_float(__float_as_uint(pl.y) + (__float_as_uint(tt.y) << 23)));
    };

    // first half of the columns, with the max of ALL 128 columns folded into the same loop
    uint32_t pk0[32];
    float2 sum_a = make_float2(0.f, 0.f);
    float mxa = -INFINITY, mxb = -INFINITY, mxc = -INFINITY, mxd = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float2 pp = ex2_pair(s[2 * i], s[2 * i + 1], neg, (i & 3) < kEmu);
        sum_a = __fadd2_rn(sum_a, pp);
        pk0[i] = pack2<kBf16>(pp.x, pp.y);
        if ((i & 3) == 0) mxa = fmaxf(mxa, fmaxf(s[4 * i], s[4 * i + 1]));
        if ((i & 3) == 0) mxb = fmaxf(mxb, fmaxf(s[4 * i + 2], s[4 * i + 3]));
        if ((i & 3) == 1) mxc = fmaxf(mxc, fmaxf(s[4 * i], s[4 * i + 1]));
        if ((i & 3) == 1) mxd = fmaxf(mxd, fmaxf(s[4 * i + 2], s[4 * i + 3]));
        if ((i & 3) == 2) mxa = fmaxf(mxa, fmaxf(s[4 * i], s[4 * i + 1]));
        if ((i & 3) == 2) mxb = fmaxf(mxb, fmaxf(s[4 * i + 2], s[4 * i + 3]));
        if ((i & 3) == 3) mxc = fmaxf(mxc, fmaxf(s[4 * i], s[4 * i + 1]));
        if ((i & 3) == 3) mxd = fmaxf(mxd, fmaxf(s[4 * i + 2], s[4 * i + 3]));
    }
    {
        const float mx = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd));
        const bool need = (mx - m_ref) * c2 > kRescaleThreshold;   // moved by more than 2^8 (never on the first tile)
        if (__any_sync(0xffffffffu, need)) {
            // slow path: rescale O_t and the running sum, redo the first half with the new reference
            float alpha = 1.f;
            if (need) {
                alpha = fast_exp2((m_ref - mx) * c2);
                m_ref = mx;
                l_run *= alpha;
                neg = -m_ref * c2;
            }
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
                uint32_t o[32];
                tmem_ld32(tO + c * 32, o);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st32(tO + c * 32, o);
            }
            sum_a = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float2 pp = ex2_pair(s[2 * i], s[2 * i + 1], neg, (i & 3) < kEmu);
                sum_a = __fadd2_rn(sum_a, pp);
                pk0[i] = pack2<kBf16>(pp.x, pp.y);
            }
        }
    }
    tmem_st32(tS, pk0);
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(bar_p0);
    // second half
    {
        uint32_t pk1[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float2 pp = ex2_pair(s[64 + 2 * i], s[64 + 2 * i + 1], neg, (i & 3) < kEmu);
            sum_a = __fadd2_rn(sum_a, pp);
            pk1[i] = pack2<kBf16>(pp.x, pp.y);
        }
        tmem_st32(tS + 32, pk1);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar_p1);
    }
    l_run += sum_a.x + sum_a.y;
}

// ---- persistent-kernel work list ----------------------------------------------------------------------
struct TileSched {
    int num_mblk;      // 256-row query blocks per (batch, head)
    int bh;            // batch * heads
    int group;         // (batch, head) pairs per L2 group
    int total;         // num_mblk * bh
};

struct WorkItem {
    int mblk, bidh, bidb;
};

FA_DEVICE WorkItem decode_item(const TileSched& ts, int n, int h, bool causal) {
    const int per_group = ts.group * ts.num_mblk;
    const int g = n / per_group;
    int r = n - g * per_group;
    const int heads_here = min(ts.group, ts.bh - g * ts.group);   // the last group may be smaller
    const int level = r / heads_here;
    const int head_local = r - level * heads_here;
    const int bhi = g * ts.group + head_local;
    WorkItem w;
    w.mblk = causal ? (ts.num_mblk - 1 - level) : level;          // heaviest causal row blocks first
    w.bidb = bhi / h;
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

// shared memory of the persistent kernel = the non-persistent layout + one 16-bit O tile used as the source of the
// TMA store (D = 128: 64 KB Q + 128 KB K/V ring + 32 KB staging = 224 KB, just inside the 227 KB limit)
template <int D> struct FwdSmemP : FwdSmem<D> {
    static constexpr int kOffStage = FwdSmem<D>::kOffKV + FwdSmem<D>::kKvStages * FwdSmem<D>::kTile;
    static constexpr int kOffBarP = kOffStage + FwdSmem<D>::kTile;
    static constexpr int kBytesP = kOffBarP + 512 + 1024;
};

// static work list of the persistent kernels (host side): see the comment at decode_item
inline TileSched make_tile_sched(const fa_fwd_params* p) {
    TileSched ts;
    ts.num_mblk = (int)((p->seqlen_q + 2 * kBlockM - 1) / (2 * kBlockM));
    ts.bh = (int)(p->b * p->h);
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
    ts.total = ts.num_mblk * ts.bh;
    return ts;
}

// launchers of the individual kernels (one translation unit each); return FA_OK / FA_ERR_*
template <int D, bool kBf16, int kEmu>
int launch_fwd_persistent(const fa_fwd_params* p, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, FwdParams kp,
                          cudaStream_t stream);
// head_dim 128 only: four softmax warpgroups (two per query tile, each thread owns half a score row)
template <bool kBf16, int kEmu>
int launch_fwd_p4(const fa_fwd_params* p, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, FwdParams kp,
                  cudaStream_t stream);

}  // namespace fa100
