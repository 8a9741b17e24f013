// flash_bwd_sm100.cu — attention backward (dQ, dK, dV) for sm_100a.
//
// Reference semantics: /root/reference/csrc/flash_attn/src/flash_bwd_preprocess_kernel.h:23-96 (D = rowsum(dO*O)),
// flash_bwd_kernel.h:28-838 (dQ) and :842-1676 (dK/dV): recompute P = exp(S*scale - LSE), dP = dO V^T,
// dS = P * (dP - D); dQ = scale * dS K; dK = scale * dS^T Q; dV = P^T dO.  Deterministic (no atomics).
// GQA: the reference writes per-q-head dK/dV into h-expanded buffers and reduces with torch::sum_out
// (flash_api.cpp:265-312); here the group sum happens inside the dK/dV kernel.
//
// Kernels in this file:
//   flash_bwd_dot_do_o_kernel_sm100   D[b,h,i] = sum_d dO*O            (HBM-bound, 16-byte loads)
//   flash_bwd_dq_kernel_sm100_rows    general-shape dQ   (one warp per query row)
//   flash_bwd_dk_dv_kernel_sm100_rows general-shape dK/dV (one warp per key row, group-summed)
// The *_rows kernels are the any-shape correctness path (ragged lengths, varlen, d=64).
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
    if (p.cu_q) { q_row0 = p.cu_q[bidb]; sq_b = p.cu_q[bidb + 1] - q_row0; }
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

// ------------------------------------------------------------------------------------------------
// general-shape dQ: one warp per (batch, head, query row); lanes split the head dimension.
// ------------------------------------------------------------------------------------------------
template <int D, bool kBf16>
__global__ void __launch_bounds__(256)
flash_bwd_dq_kernel_sm100_rows(const BwdParams p) {
    constexpr int E = D / 32;  // elements per lane (2 or 4)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bidb = blockIdx.z, bidh = blockIdx.y, bidh_k = bidh / p.hratio;
    int q_row0 = 0, k_row0 = 0, sq_b = p.sq, sk_b = p.sk;
    if (p.cu_q) {
        q_row0 = p.cu_q[bidb]; sq_b = p.cu_q[bidb + 1] - q_row0;
        k_row0 = p.cu_k[bidb]; sk_b = p.cu_k[bidb + 1] - k_row0;
    }
    const int i = blockIdx.x * 8 + warp;
    if (i >= sq_b) return;
    const int64_t qrow = (p.cu_q ? (int64_t)q_row0 : (int64_t)bidb * p.sq) + i;
    const int64_t krow0 = p.cu_q ? (int64_t)k_row0 : (int64_t)bidb * p.sk;
    const uint16_t* q16 = reinterpret_cast<const uint16_t*>(p.q) + (qrow * p.h + bidh) * D + lane * E;
    const uint16_t* g16 = reinterpret_cast<const uint16_t*>(p.dout) + (qrow * p.h + bidh) * D + lane * E;
    float qv[E], gv[E], acc[E];
#pragma unroll
    for (int e = 0; e < E; e += 2) {
        const float2 a = unpack2<kBf16>(*reinterpret_cast<const uint32_t*>(q16 + e));
        const float2 g = unpack2<kBf16>(*reinterpret_cast<const uint32_t*>(g16 + e));
        qv[e] = a.x; qv[e + 1] = a.y; gv[e] = g.x; gv[e + 1] = g.y; acc[e] = 0.f; acc[e + 1] = 0.f;
    }
    const int64_t stat = ((int64_t)bidb * p.h + bidh) * p.sq + i;
    const float lse = p.lse[stat], dsum = p.dsum[stat];
    int jmax = sk_b;
    if (p.is_causal) jmax = min(sk_b, max(0, i + sk_b - sq_b + 1));
    for (int j = 0; j < jmax; ++j) {
        const uint16_t* k16 = reinterpret_cast<const uint16_t*>(p.k) + ((krow0 + j) * p.h_k + bidh_k) * D + lane * E;
        const uint16_t* v16 = reinterpret_cast<const uint16_t*>(p.v) + ((krow0 + j) * p.h_k + bidh_k) * D + lane * E;
        float kv[E], s = 0.f, dp = 0.f;
#pragma unroll
        for (int e = 0; e < E; e += 2) {
            const float2 a = unpack2<kBf16>(*reinterpret_cast<const uint32_t*>(k16 + e));
            const float2 c = unpack2<kBf16>(*reinterpret_cast<const uint32_t*>(v16 + e));
            kv[e] = a.x; kv[e + 1] = a.y;
            s = fmaf(qv[e], a.x, s); s = fmaf(qv[e + 1], a.y, s);
            dp = fmaf(gv[e], c.x, dp); dp = fmaf(gv[e + 1], c.y, dp);
        }
        s = warp_sum(s); dp = warp_sum(dp);
        const float pr = __expf(s * p.scale - lse);
        const float ds = pr * (dp - dsum);
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = fmaf(ds, kv[e], acc[e]);
    }
    uint16_t* out = reinterpret_cast<uint16_t*>(p.dq) + (qrow * p.h + bidh) * D + lane * E;
#pragma unroll
    for (int e = 0; e < E; e += 2)
        *reinterpret_cast<uint32_t*>(out + e) = pack2<kBf16>(acc[e] * p.scale, acc[e + 1] * p.scale);
}

// ------------------------------------------------------------------------------------------------
// general-shape dK/dV: one warp per (batch, kv head, key row); sums over the GQA group in-kernel.
// ------------------------------------------------------------------------------------------------
template <int D, bool kBf16>
__global__ void __launch_bounds__(256)
flash_bwd_dk_dv_kernel_sm100_rows(const BwdParams p) {
    constexpr int E = D / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bidb = blockIdx.z, bidh_k = blockIdx.y;
    int q_row0 = 0, k_row0 = 0, sq_b = p.sq, sk_b = p.sk;
    if (p.cu_q) {
        q_row0 = p.cu_q[bidb]; sq_b = p.cu_q[bidb + 1] - q_row0;
        k_row0 = p.cu_k[bidb]; sk_b = p.cu_k[bidb + 1] - k_row0;
    }
    const int j = blockIdx.x * 8 + warp;
    if (j >= sk_b) return;
    const int64_t qrow0 = p.cu_q ? (int64_t)q_row0 : (int64_t)bidb * p.sq;
    const int64_t krow = (p.cu_q ? (int64_t)k_row0 : (int64_t)bidb * p.sk) + j;
    const uint16_t* k16 = reinterpret_cast<const uint16_t*>(p.k) + (krow * p.h_k + bidh_k) * D + lane * E;
    const uint16_t* v16 = reinterpret_cast<const uint16_t*>(p.v) + (krow * p.h_k + bidh_k) * D + lane * E;
    float kv[E], vv[E], dk[E], dv[E];
#pragma unroll
    for (int e = 0; e < E; e += 2) {
        const float2 a = unpack2<kBf16>(*reinterpret_cast<const uint32_t*>(k16 + e));
        const float2 c = unpack2<kBf16>(*reinterpret_cast<const uint32_t*>(v16 + e));
        kv[e] = a.x; kv[e + 1] = a.y; vv[e] = c.x; vv[e + 1] = c.y;
        dk[e] = dk[e + 1] = dv[e] = dv[e + 1] = 0.f;
    }
    int imin = 0;
    if (p.is_causal) imin = max(0, j - (sk_b - sq_b));  // visible iff j <= i + sk - sq
    for (int g = 0; g < p.hratio; ++g) {
        const int bidh = bidh_k * p.hratio + g;
        const float* lse = p.lse + ((int64_t)bidb * p.h + bidh) * p.sq;
        const float* dsum = p.dsum + ((int64_t)bidb * p.h + bidh) * p.sq;
        for (int i = imin; i < sq_b; ++i) {
            const uint16_t* q16 = reinterpret_cast<const uint16_t*>(p.q) + ((qrow0 + i) * p.h + bidh) * D + lane * E;
            const uint16_t* g16 = reinterpret_cast<const uint16_t*>(p.dout) + ((qrow0 + i) * p.h + bidh) * D + lane * E;
            float qv[E], gv[E], s = 0.f, dp = 0.f;
#pragma unroll
            for (int e = 0; e < E; e += 2) {
                const float2 a = unpack2<kBf16>(*reinterpret_cast<const uint32_t*>(q16 + e));
                const float2 c = unpack2<kBf16>(*reinterpret_cast<const uint32_t*>(g16 + e));
                qv[e] = a.x; qv[e + 1] = a.y; gv[e] = c.x; gv[e + 1] = c.y;
                s = fmaf(a.x, kv[e], s); s = fmaf(a.y, kv[e + 1], s);
                dp = fmaf(c.x, vv[e], dp); dp = fmaf(c.y, vv[e + 1], dp);
            }
            s = warp_sum(s); dp = warp_sum(dp);
            const float pr = __expf(s * p.scale - lse[i]);
            const float ds = pr * (dp - dsum[i]);
#pragma unroll
            for (int e = 0; e < E; ++e) { dv[e] = fmaf(pr, gv[e], dv[e]); dk[e] = fmaf(ds, qv[e], dk[e]); }
        }
    }
    uint16_t* okk = reinterpret_cast<uint16_t*>(p.dk) + (krow * p.h_k + bidh_k) * D + lane * E;
    uint16_t* ovv = reinterpret_cast<uint16_t*>(p.dv) + (krow * p.h_k + bidh_k) * D + lane * E;
#pragma unroll
    for (int e = 0; e < E; e += 2) {
        *reinterpret_cast<uint32_t*>(okk + e) = pack2<kBf16>(dk[e] * p.scale, dk[e + 1] * p.scale);
        *reinterpret_cast<uint32_t*>(ovv + e) = pack2<kBf16>(dv[e], dv[e + 1]);
    }
}

static bool use_row_kernels() {
    // FA_B200_BWD=rows selects the CUDA-core any-shape kernels (debug / cross-check); default is the tcgen05 path
    // (fused dQ/dK/dV kernel for head_dim 128, "det" = the two deterministic tcgen05 kernels)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FA_B200_BWD");
        v = (e && e[0] == 'r') ? 1 : 0;
    }
    return v == 1;
}

template <int D, bool kBf16> static int launch_bwd_rows(const BwdParams& kp, bool bf16, cudaStream_t stream) {
    constexpr int kRowsPerBlockD = 8 * (32 / (D / 8));
    if (kp.sq > 0) {
        dim3 g1((kp.sq + kRowsPerBlockD - 1) / kRowsPerBlockD, kp.h, kp.b);
        flash_bwd_dot_do_o_kernel_sm100<D, kBf16><<<g1, 256, 0, stream>>>(kp);
        FA_CUDA_CHECK(cudaGetLastError());
        count_launch();
    }
    if (!use_row_kernels()) {
        const int rc = launch_bwd_tc_sm100(kp, bf16, stream);
        if (rc >= 0) return rc;   // < 0: degenerate sizes, fall through to the row kernels (they write the zeros)
    }
    if (kp.sq > 0) {
        dim3 g2((kp.sq + 7) / 8, kp.h, kp.b);
        flash_bwd_dq_kernel_sm100_rows<D, kBf16><<<g2, 256, 0, stream>>>(kp);
        FA_CUDA_CHECK(cudaGetLastError());
        count_launch();
    }
    if (kp.sk > 0) {
        dim3 g3((kp.sk + 7) / 8, kp.h_k, kp.b);
        flash_bwd_dk_dv_kernel_sm100_rows<D, kBf16><<<g3, 256, 0, stream>>>(kp);
        FA_CUDA_CHECK(cudaGetLastError());
        count_launch();
    }
    return FA_OK;
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
    kp.dqacc = static_cast<float*>(p->workspace);
    kp.sq_pad = (kp.sq + 63) / 64 * 64;
    const bool bf16 = f->dtype == FA_DTYPE_BF16;
    if (f->d == 128) return bf16 ? launch_bwd_rows<128, true>(kp, bf16, stream) : launch_bwd_rows<128, false>(kp, bf16, stream);
    if (f->d == 64) return bf16 ? launch_bwd_rows<64, true>(kp, bf16, stream) : launch_bwd_rows<64, false>(kp, bf16, stream);
    set_error("head_dim %lld not supported (64 or 128)", (long long)f->d);
    return FA_ERR_INVALID_ARG;
}

}  // namespace fa100
