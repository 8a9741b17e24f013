// flash_bwd_sm100.cu — attention backward (dQ, dK, dV) for sm_100a.
//
// Reference semantics: /root/reference/csrc/flash_attn/src/flash_bwd_preprocess_kernel.h:23-96 (D = rowsum(dO*O)),
// flash_bwd_kernel.h:28-838 (dQ) and :842-1676 (dK/dV): recompute P = exp(S*scale - LSE), dP = dO V^T,
// dS = P * (dP - D); dQ = scale * dS K; dK = scale * dS^T Q; dV = P^T dO.
// Determinism: dK and dV are bit-reproducible on every path.  dQ is bit-reproducible on the two-kernel path (no
// workspace, or FA_B200_BWD=det) like the reference; the default fused path (head_dim 128 and 64) accumulates dQ with fp32
// reductions (red.global.add / cp.reduce.async.bulk) whose order varies from run to run (last-bit differences).
// GQA: the reference writes per-q-head dK/dV into h-expanded buffers and reduces with torch::sum_out
// (flash_api.cpp:265-312); here the group sum happens inside the dK/dV kernel.
//
// Kernels in this file:
//   flash_bwd_dot_do_o_kernel_sm100   D[b,h,i] = sum_d dO*O            (HBM-bound, 16-byte loads)
// The dQ / dK / dV kernels (tcgen05) live in flash_bwd_tc_sm100.cu.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "fa_common.h"
#include "sm100_ptx.cuh"
#include "flash_bwd_params.h"

namespace fa100 {


template <bool kBf16> FA_DEVICE float2 unpack2(uint32_t w) {
    if constexpr (kBf16) {
        return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    } else {
        const __half2 h = *reinterpret_cast<const __half2*>(&w);
        return __half22float2(h);
    }
}

FA_DEVICE float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// ------------------------------------------------------------------------------------------------
// D = rowsum(dO * O): a half-warp (d=128) or quarter-warp (d=64) per row, one 16-byte load per lane.
// ------------------------------------------------------------------------------------------------
template <int D, bool kBf16>
__global__ void __launch_bounds__(256)
flash_bwd_dot_do_o_kernel_sm100(const BwdParams p) {
    constexpr int kLanesPerRow = D / 8;            // 16 or 8
    constexpr int kRowsPerWarp = 32 / kLanesPerRow;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bidb = blockIdx.z, bidh = blockIdx.y;
    int q_row0 = 0, sq_b = p.sq;
    if (p.cu_q) { q_row0 = p.cu_q[bidb]; sq_b = min(p.cu_q[bidb + 1] - q_row0, p.sq); }
    const int64_t row_base = p.cu_q ? (int64_t)q_row0 : (int64_t)bidb * p.sq;
    const int row = (blockIdx.x * 8 + warp) * kRowsPerWarp + lane / kLanesPerRow;
    float acc = 0.f;
    if (row < sq_b) {
        const int64_t off = ((row_base + row) * p.h + bidh) * D + (lane % kLanesPerRow) * 8;
        const uint4 a = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.o) + off);
        const uint4 g = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.dout) + off);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 x = unpack2<kBf16>(aw[i]), y = unpack2<kBf16>(gw[i]);
            acc = fmaf(x.x, y.x, acc);
            acc = fmaf(x.y, y.y, acc);
        }
    }
#pragma unroll
    for (int o = kLanesPerRow / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (row < sq_b && (lane % kLanesPerRow) == 0) p.dsum[((int64_t)bidb * p.h + bidh) * p.sq + row] = acc;
}

template <int D, bool kBf16> static int launch_bwd(const BwdParams& kp, bool bf16, cudaStream_t stream) {
    constexpr int kRowsPerBlockD = 8 * (32 / (D / 8));
    const int64_t rows_q = kp.cu_q ? kp.total_q : (int64_t)kp.b * kp.sq;
    const int64_t rows_k = kp.cu_q ? kp.total_k : (int64_t)kp.b * kp.sk;
    if (rows_q == 0 || rows_k == 0 || kp.sq == 0 || kp.sk == 0) {
        // degenerate: no (query, key) pair exists, every gradient that has elements is zero
        if (rows_q > 0) FA_CUDA_CHECK(cudaMemsetAsync(kp.dq, 0, (size_t)rows_q * kp.h * D * 2, stream));
        if (rows_k > 0) {
            FA_CUDA_CHECK(cudaMemsetAsync(kp.dk, 0, (size_t)rows_k * kp.h_k * D * 2, stream));
            FA_CUDA_CHECK(cudaMemsetAsync(kp.dv, 0, (size_t)rows_k * kp.h_k * D * 2, stream));
        }
        return FA_OK;
    }
    dim3 g1((kp.sq + kRowsPerBlockD - 1) / kRowsPerBlockD, kp.h, kp.b);
    flash_bwd_dot_do_o_kernel_sm100<D, kBf16><<<g1, 256, 0, stream>>>(kp);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    const int rc = launch_bwd_tc_sm100(kp, bf16, stream);
    if (rc < 0) { set_error("backward: unsupported shape"); return FA_ERR_INVALID_ARG; }
    return rc;
}

int launch_bwd_sm100(const fa_bwd_params* p, cudaStream_t stream) {
    const fa_fwd_params* f = &p->fwd;
    if (f->b == 0 || f->h == 0) return FA_OK;
    BwdParams kp;
    kp.q = f->q; kp.k = f->k; kp.v = f->v; kp.o = f->o; kp.dout = p->dout; kp.lse = f->lse; kp.dsum = p->dsum;
    kp.dq = p->dq; kp.dk = p->dk; kp.dv = p->dv; kp.cu_q = f->cu_seqlens_q; kp.cu_k = f->cu_seqlens_k;
    kp.b = (int)f->b; kp.sq = (int)f->seqlen_q; kp.sk = (int)f->seqlen_k; kp.h = (int)f->h; kp.h_k = (int)f->h_k;
    kp.hratio = (int)(f->h / f->h_k); kp.d = (int)f->d; kp.is_causal = f->is_causal;
    kp.scale = 1.0f / sqrtf((float)f->d);
    kp.total_q = f->total_q; kp.total_k = f->total_k; kp.trace = nullptr;
#ifdef FA_TRACE
    kp.trace = fa_trace_buffer(stream);   // clock64 build (make trace): scripts/trace_bwd.py reads it back
#endif
    kp.dqacc = static_cast<float*>(p->workspace);
    kp.sq_pad = (kp.sq + 63) / 64 * 64;
    const bool bf16 = f->dtype == FA_DTYPE_BF16;
    if (f->d == 128) return bf16 ? launch_bwd<128, true>(kp, bf16, stream) : launch_bwd<128, false>(kp, bf16, stream);
    if (f->d == 64) return bf16 ? launch_bwd<64, true>(kp, bf16, stream) : launch_bwd<64, false>(kp, bf16, stream);
    set_error("head_dim %lld not supported (64 or 128)", (long long)f->d);
    return FA_ERR_INVALID_ARG;
}

}  // namespace fa100
