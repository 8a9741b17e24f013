// flash_fwd_sm100.cu — host side of the Blackwell (sm_100a) fused attention forward.
//
//   O = softmax(Q K^T / sqrt(d) + mask) V,   LSE = ln sum exp(.)      (reference semantics:
//   /root/reference/csrc/flash_attn/src/flash_fwd_kernel.h:23-789, mask.h:20-72)
//
// Replaces the reference's run_mha_fwd_<Headdim,Is_causal> launch templates
// (/root/reference/csrc/flash_attn/src/flash_fwd_launch_template.h:45-111): builds the TMA views of Q, K, V
// ([batch][row][head][d] boxes of 128 rows x 64 columns, 128-byte swizzle; a packed varlen tensor is one "batch" of
// total rows) and launches the one forward kernel of this library, flash_fwd_kernel_sm100_p4 (flash_fwd_p4_sm100.cu),
// for head_dim 64 or 128, fp16 or bf16.  Nothing here resembles the reference's ld.global -> st.shared -> ldmatrix ->
// mma.sync pipeline: tiles arrive by TMA, both contractions run on tcgen05 with accumulators in TMEM.
#include <stdlib.h>

#include "flash_fwd_common.cuh"

namespace fa100 {

static int fwd_emu(int d) {
    // FA_B200_EMU: share of the exponentials evaluated by a degree-3 polynomial on the FMA pipe instead of MUFU.EX2
    //   0 = none, 4 = 1 pair in 8, 1 = 2 in 8, 3 = 3 in 8, 2 = 4 in 8.   (tuning knob; defaults measured on B200)
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("FA_B200_EMU");
        v = e ? atoi(e) : -1;
        if (v < -1 || v > 4) v = -1;
    }
    if (v >= 0) return v;
    (void)d;
    // 1 pair in 8 through the polynomial: with one MMA-issuing warp per tile the softmax phase is on the critical path again and
    // taking an eighth of the exponentials off the MUFU pipe pays +2.3 % (C2) / +2.4 % (C3) / +2.8 % (head_dim 64) in a
    // sustained loop; 2 in 8 measures the same, 3 in 8 nothing (profiles/r02_run19.log)
    return 4;
}

template <int D>
static int dispatch_p4(const fa_fwd_params* p, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const FwdParams& kp, cudaStream_t stream) {
    const bool bf16 = p->dtype == FA_DTYPE_BF16;
    switch (fwd_emu(D)) {
        case 0: return bf16 ? launch_fwd_p4<D, true, 0>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<D, false, 0>(p, tq, tk, tv, kp, stream);
        case 1: return bf16 ? launch_fwd_p4<D, true, 1>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<D, false, 1>(p, tq, tk, tv, kp, stream);
        case 3: return bf16 ? launch_fwd_p4<D, true, 3>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<D, false, 3>(p, tq, tk, tv, kp, stream);
        case 4: return bf16 ? launch_fwd_p4<D, true, 4>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<D, false, 4>(p, tq, tk, tv, kp, stream);
        default: return bf16 ? launch_fwd_p4<D, true, 2>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<D, false, 2>(p, tq, tk, tv, kp, stream);
    }
}

int launch_fwd_sm100(const fa_fwd_params* p, cudaStream_t stream) {
    const bool varlen = p->cu_seqlens_q != nullptr;
    const bool bf16 = p->dtype == FA_DTYPE_BF16;
    const uint64_t D = (uint64_t)p->d;
    if (p->b == 0 || p->seqlen_q == 0 || p->h == 0) return FA_OK;
    if (varlen ? (p->total_q == 0) : false) return FA_OK;

    FwdParams kp;
    kp.o = p->o; kp.lse = p->lse; kp.cu_q = p->cu_seqlens_q; kp.cu_k = p->cu_seqlens_k;
    kp.b = (int)p->b; kp.sq = (int)p->seqlen_q; kp.sk = (int)p->seqlen_k; kp.h = (int)p->h; kp.h_k = (int)p->h_k;
    kp.hratio = (int)(p->h / p->h_k);
    kp.is_causal = p->is_causal;
    static int exact = -1;   // FA_B200_FWD_EXACT=1: exact row max at every key step (the retry path of the default kernel) for A/B runs
    if (exact < 0) { const char* e = getenv("FA_B200_FWD_EXACT"); exact = (e && atoi(e) != 0) ? 1 : 0; }
    kp.exact = exact;
    kp.scale = 1.0f / sqrtf((float)p->d);
    kp.scale_log2 = kp.scale * 1.4426950408889634f;
    kp.inv_scale_log2 = 1.0f / kp.scale_log2;
    kp.trace = nullptr;
#ifdef FA_TRACE
    kp.trace = fa_trace_buffer(stream);   // clock64 build (make trace): scripts/trace_fwd.py reads it back
#endif

    // TMA views: [batch][row][head][d]; packed varlen tensors are one "batch" of total rows.
    const uint64_t rows_q = varlen ? (uint64_t)p->total_q : (uint64_t)p->seqlen_q;
    const uint64_t rows_k = varlen ? (uint64_t)p->total_k : (uint64_t)p->seqlen_k;
    const uint64_t nb = varlen ? 1 : (uint64_t)p->b;
    CUtensorMap tq, tk, tv;
    const uint32_t box[4] = {64, 1, (uint32_t)kBlockM, 1};
    {
        const uint64_t dims[4] = {D, (uint64_t)p->h, rows_q, nb};
        const uint64_t str[3] = {D * 2, (uint64_t)p->h * D * 2, rows_q * (uint64_t)p->h * D * 2};
        int rc = encode_tmap_4d(&tq, p->q, bf16, dims, str, box);
        if (rc != FA_OK) return rc;
    }
    if (rows_k > 0) {
        const uint64_t dims[4] = {D, (uint64_t)p->h_k, rows_k, nb};
        const uint64_t str[3] = {D * 2, (uint64_t)p->h_k * D * 2, rows_k * (uint64_t)p->h_k * D * 2};
        int rc = encode_tmap_4d(&tk, p->k, bf16, dims, str, box);
        if (rc != FA_OK) return rc;
        rc = encode_tmap_4d(&tv, p->v, bf16, dims, str, box);
        if (rc != FA_OK) return rc;
    } else {
        tk = tq; tv = tq;  // never dereferenced: every tile has n_blocks == 0
    }
    if (p->d == 128) return dispatch_p4<128>(p, tq, tk, tv, kp, stream);
    if (p->d == 64) return dispatch_p4<64>(p, tq, tk, tv, kp, stream);
    set_error("head_dim %lld not supported (64 or 128)", (long long)p->d);
    return FA_ERR_INVALID_ARG;
}

}  // namespace fa100
