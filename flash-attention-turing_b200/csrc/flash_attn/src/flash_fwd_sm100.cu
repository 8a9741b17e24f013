// flash_fwd_sm100.cu — Blackwell (sm_100a) fused attention forward.
//
//   O = softmax(Q K^T / sqrt(d) + mask) V,   LSE = ln sum exp(.)      (reference semantics:
//   /root/reference/csrc/flash_attn/src/flash_fwd_kernel.h:23-789, mask.h:20-72)
//
// Data path (nothing here resembles the reference's ld.global -> st.shared -> ldmatrix -> mma.sync
// pipeline; it is designed for the B200 SM):
//   * Q, K, V tiles are fetched by TMA (cp.async.bulk.tensor, 128-byte swizzle) straight into shared
//     memory; out-of-range rows are zero-filled by the TMA unit, so ragged lengths need no special
//     load path.
//   * S = Q K^T and O += P V run on the 5th-gen tensor cores (tcgen05.mma, kind::f16) issued by one
//     thread; S and O live in tensor memory (TMEM), P is written back to TMEM as the A operand of PV
//     (V is consumed in place as an MN-major B operand — no transpose anywhere).
//   * The online softmax runs with one thread per query row (tcgen05.ld 32x32b), so row max / row sum
//     need no cross-thread reduction at all.
//
// This file holds the general kernel "flash_fwd_kernel_sm100_g": one 128-row Q tile per CTA, any
// sequence lengths, causal, GQA, varlen, d in {64,128}, fp16/bf16.
#include <stdlib.h>

#include "flash_fwd_common.cuh"

namespace fa100 {

template <int D> struct FwdSmemG {
    static constexpr int kSlab = kBlockM * 128;             // one 64-column slab of a 128-row tile (16 KB)
    static constexpr int kTile = kBlockM * D * 2;           // one full tile
    static constexpr int kStages = 2;
    static constexpr int kOffQ = 0;
    static constexpr int kOffK = kTile;
    static constexpr int kOffV = kOffK + kStages * kTile;
    static constexpr int kOffBar = kOffV + kStages * kTile;
    static constexpr int kBytes = kOffBar + 256 + 1024;     // + barriers + alignment slack
};

// TMEM column map (512 columns allocated; 1 CTA per SM)
constexpr uint32_t kTmemS = 0;     // 128 fp32 columns
constexpr uint32_t kTmemP = 128;   // 64 columns (128 x 16-bit)
constexpr uint32_t kTmemO = 256;   // D fp32 columns

template <int D, bool kBf16>
__global__ void __launch_bounds__(192, 1)
flash_fwd_kernel_sm100_g(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
    using L = FwdSmemG<D>;
    constexpr int kSlabs = D / 64;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int m0 = blockIdx.x * kBlockM;
    const int bidh = blockIdx.y;
    const int bidb = blockIdx.z;
    const int bidh_k = bidh / p.hratio;

    // ---- per-sequence geometry (BlockInfo equivalent, block_info.h:3-27, with 64-bit row math) ----
    int q_row0, k_row0, sq_b, sk_b, tma_b;
    if (p.cu_q != nullptr) {
        q_row0 = p.cu_q[bidb];
        sq_b = p.cu_q[bidb + 1] - q_row0;
        k_row0 = p.cu_k[bidb];
        sk_b = p.cu_k[bidb + 1] - k_row0;
        tma_b = 0;
    } else {
        q_row0 = 0; k_row0 = 0; sq_b = p.sq; sk_b = p.sk; tma_b = bidb;
    }
    if (m0 >= sq_b) return;
    const int causal_off = sk_b - sq_b;  // keep (i,j) iff j <= i + causal_off
    int kv_end = sk_b;
    if (p.is_causal) kv_end = min(sk_b, max(0, m0 + kBlockM + causal_off));
    const int n_blocks = (kv_end + kBlockN - 1) / kBlockN;

    const int64_t o_row_base = (p.cu_q != nullptr) ? (int64_t)q_row0 : (int64_t)bidb * p.sq;
    float* lse_row = p.lse + ((int64_t)bidb * p.h + bidh) * p.sq;

    if (n_blocks == 0) {
        // every row of this tile is fully masked: O = 0, LSE = 0 (flash_fwd_kernel.h:275, :717-730, :766-785)
        for (int idx = tid; idx < kBlockM * (D / 8); idx += blockDim.x) {
            const int r = idx / (D / 8), c = idx % (D / 8);
            if (m0 + r < sq_b) {
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.o) +
                                                      ((o_row_base + m0 + r) * p.h + bidh) * D) + c;
                *dst = make_uint4(0, 0, 0, 0);
            }
        }
        if (tid < kBlockM && m0 + tid < sq_b) lse_row[m0 + tid] = 0.f;
        return;
    }

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + L::kOffQ;
    uint8_t* sK = smem + L::kOffK;
    uint8_t* sV = smem + L::kOffV;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* bar_q = bars + 0;
    uint64_t* bar_kv_full = bars + 1;   // [2]
    uint64_t* bar_kv_empty = bars + 3;  // [2]
    uint64_t* bar_s_full = bars + 5;
    uint64_t* bar_s_empty = bars + 6;
    uint64_t* bar_p_full = bars + 7;
    uint64_t* bar_o_done = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    if (warp == 4) {
        if (lane == 0) {
            mbar_init(bar_q, 1);
            mbar_init(&bar_kv_full[0], 1); mbar_init(&bar_kv_full[1], 1);
            mbar_init(&bar_kv_empty[0], 1); mbar_init(&bar_kv_empty[1], 1);
            mbar_init(bar_s_full, 1);
            mbar_init(bar_s_empty, kBlockM);
            mbar_init(bar_p_full, kBlockM);
            mbar_init(bar_o_done, 1);
            fence_barrier_init();
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_q, L::kTile);
            for (int s = 0; s < kSlabs; ++s)
                tma_load_4d(sQ + s * L::kSlab, &tmQ, bar_q, s * 64, bidh, q_row0 + m0, tma_b);
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j & 1;
                mbar_wait(&bar_kv_empty[st], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar_kv_full[st], 2 * L::kTile);
                for (int s = 0; s < kSlabs; ++s)
                    tma_load_4d(sK + st * L::kTile + s * L::kSlab, &tmK, &bar_kv_full[st], s * 64, bidh_k,
                                k_row0 + j * kBlockN, tma_b);
                for (int s = 0; s < kSlabs; ++s)
                    tma_load_4d(sV + st * L::kTile + s * L::kSlab, &tmV, &bar_kv_full[st], s * 64, bidh_k,
                                k_row0 + j * kBlockN, tma_b);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc(kBf16, kBlockM, kBlockN, false, false);
            constexpr uint32_t idesc_pv = make_idesc(kBf16, kBlockM, D, false, true);
            const uint32_t q_addr = smem_u32(sQ);
            mbar_wait(bar_q, 0);
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j & 1;
                const uint32_t k_addr = smem_u32(sK + st * L::kTile);
                const uint32_t v_addr = smem_u32(sV + st * L::kTile);
                mbar_wait(&bar_kv_full[st], (j >> 1) & 1);
                mbar_wait(bar_s_empty, (j & 1) ^ 1);
                tc_fence_after();
                // S[128 x 128] = Q[128 x D] K[128 x D]^T : both K-major, 16 elements (32 B) per MMA step
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * L::kSlab + (kk & 3) * 32;
                    umma_ss(tmem_base + kTmemS, make_smem_desc(q_addr + off, 16, 1024),
                            make_smem_desc(k_addr + off, 16, 1024), idesc_s, kk > 0);
                }
                tc_commit(bar_s_full);
                mbar_wait(bar_p_full, j & 1);
                tc_fence_after();
                // O[128 x D] += P[128 x 128] V[128 x D] : P from TMEM, V MN-major (16 key rows = 2048 B per step)
#pragma unroll
                for (int kk = 0; kk < kBlockN / 16; ++kk) {
                    umma_ts(tmem_base + kTmemO, tmem_base + kTmemP + kk * 8,
                            make_smem_desc(v_addr + kk * 2048, L::kSlab, 1024), idesc_pv, (j > 0 || kk > 0));
                }
                tc_commit(&bar_kv_empty[st]);
                tc_commit(bar_o_done);
            }
        }
    } else {
        // ===================== softmax / correction / epilogue: one thread per query row =====================
        const int row = m0 + tid;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const uint32_t tS = tmem_base + lane_base + kTmemS;
        const uint32_t tP = tmem_base + lane_base + kTmemP;
        const uint32_t tO = tmem_base + lane_base + kTmemO;
        // last visible key column for this row
        int col_limit = sk_b - 1;
        if (p.is_causal) col_limit = min(col_limit, row + causal_off);
        float m_run = -INFINITY, l_run = 0.f;
        const float c2 = p.scale_log2;

        for (int j = 0; j < n_blocks; ++j) {
            const int n0 = j * kBlockN;
            mbar_wait(bar_s_full, j & 1);
            tc_fence_after();
            float s[kBlockN];
#pragma unroll
            for (int c = 0; c < kBlockN / 32; ++c)
                tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(bar_s_empty);

            const bool need_mask = (n0 + kBlockN > sk_b) || (p.is_causal && (n0 + kBlockN - 1 > m0 + causal_off));
            if (need_mask) {
                const int lim = col_limit - n0;
#pragma unroll
                for (int c = 0; c < kBlockN; ++c)
                    if (c > lim) s[c] = -INFINITY;
            }
            float mx = s[0];
#pragma unroll
            for (int c = 1; c < kBlockN; ++c) mx = fmaxf(mx, s[c]);
            const float m_new = fmaxf(m_run, mx);
            const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
            const float alpha = fast_exp2((m_run - m_safe) * c2);
            const float neg = -m_safe * c2;
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < kBlockN; ++c) {
                s[c] = fast_exp2(fmaf(s[c], c2, neg));
                sum += s[c];
            }
            l_run = l_run * alpha + sum;
            m_run = m_new;

            if (j > 0) {
                mbar_wait(bar_o_done, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
                    for (int c = 0; c < D / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld32(tO + c * 32, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(tO + c * 32, o);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kBlockN / 64; ++c) {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) pk[i] = pack2<kBf16>(s[c * 64 + 2 * i], s[c * 64 + 2 * i + 1]);
                tmem_st32(tP + c * 32, pk);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar_p_full);
        }

        // ---- epilogue: O / l -> 16-bit -> smem (reuses the Q tile) -> coalesced global stores ----
        mbar_wait(bar_o_done, (n_blocks - 1) & 1);
        tc_fence_after();
        const float inv_l = (l_run > 0.f) ? (1.f / l_run) : 0.f;
        uint8_t* sO = sQ;  // all S MMAs have retired (o_done covers every earlier tcgen05.mma)
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < 4; ++g) {  // 4 x 16-byte chunks (8 values each)
                uint4 v;
                v.x = pack2<kBf16>(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
                v.y = pack2<kBf16>(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
                v.z = pack2<kBf16>(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
                v.w = pack2<kBf16>(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
                const int chunk = c * 4 + g;  // 16-byte chunk index within the row
                *reinterpret_cast<uint4*>(sO + tid * (D * 2) + ((chunk ^ (tid & 7)) * 16)) = v;
            }
        }
        if (row < sq_b) lse_row[row] = (l_run > 0.f) ? (m_run * p.scale + logf(l_run)) : 0.f;
        tc_fence_before();
        named_bar_sync(1, kBlockM);
        constexpr int kChunksPerRow = D / 8;
        uint16_t* o_base = reinterpret_cast<uint16_t*>(p.o);
#pragma unroll 4
        for (int idx = tid; idx < kBlockM * kChunksPerRow; idx += kBlockM) {
            const int r = idx / kChunksPerRow, ch = idx % kChunksPerRow;
            if (m0 + r < sq_b) {
                const uint4 v = *reinterpret_cast<const uint4*>(sO + r * (D * 2) + ((ch ^ (r & 7)) * 16));
                *(reinterpret_cast<uint4*>(o_base + ((o_row_base + m0 + r) * p.h + bidh) * D) + ch) = v;
            }
        }
    }

    __syncthreads();
    if (warp == 4) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// =================================================================================================
// flash_fwd_kernel_sm100 — the warp-specialised forward.
//
// One CTA owns TWO 128-row query tiles of one (batch, head) and walks the key/value tiles once for both:
//
//   warpgroup 0 (warps 0-3)  : softmax + output epilogue for query tile 0   (one thread per row)
//   warpgroup 1 (warps 4-7)  : softmax + output epilogue for query tile 1
//   warp 8                   : tcgen05.mma issuer (one elected thread)
//   warp 9                   : TMA producer (Q tiles, then the K/V ring)
//   warps 10-11              : idle (they only donate registers via setmaxnreg)
//
// TMEM (512 columns): S0 [0,128)  S1 [128,256)  O0 [256,256+D)  O1 [384,384+D).  P_t overwrites the first
// 64 columns of S_t (two 16-bit values per column) and is consumed from there as the A operand of P V.
// The tensor pipe executes MMAs in issue order, and the issue order is
//     S0_0 S1_0 | PV0_0 S0_1 PV1_0 S1_1 | PV0_1 S0_2 PV1_1 S1_2 | ...
// so while warpgroup t runs the softmax of S_t the tensor cores work on the other tile, and "S_t_{j+1} is
// complete" implies "P V_t_j is complete": the softmax thread may then overwrite P_t and (rarely) rescale
// O_t without any further synchronisation.
//
// Online softmax with lazy rescaling: the running reference max m_ref of a row only moves when the new tile
// max exceeds it by more than 8 in the log2 domain (P <= 2^8 stays exact enough in fp32 accumulators and in
// 16-bit P); O_t is rescaled in TMEM only on those steps.  The final O / l and LSE = m_ref*scale + ln l are
// unchanged by this (the reference rescales on every tile, flash_fwd_kernel.h:675-679).
// =================================================================================================
template <int D, bool kBf16, int kEmu>
__global__ void __launch_bounds__(384, 1)
flash_fwd_kernel_sm100(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
    using L = FwdSmem<D>;
    constexpr int kSlabs = D / 64;
    constexpr int kStages = L::kKvStages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int wg = warp >> 2;
    // heavier (later) causal row blocks first
    const int mblk = p.is_causal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
    const int m0 = mblk * (2 * kBlockM);
    const int bidh = blockIdx.y;
    const int bidb = blockIdx.z;
    const int bidh_k = bidh / p.hratio;

    int q_row0, k_row0, sq_b, sk_b, tma_b;
    if (p.cu_q != nullptr) {
        q_row0 = p.cu_q[bidb];
        sq_b = p.cu_q[bidb + 1] - q_row0;
        k_row0 = p.cu_k[bidb];
        sk_b = p.cu_k[bidb + 1] - k_row0;
        tma_b = 0;
    } else {
        q_row0 = 0; k_row0 = 0; sq_b = p.sq; sk_b = p.sk; tma_b = bidb;
    }
    if (m0 >= sq_b) return;
    const int causal_off = sk_b - sq_b;
    // key tiles needed by each query tile (0 if the tile has no rows or sees no key)
    int nblk[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int mt = m0 + t * kBlockM;
        int kv_end = (mt < sq_b) ? sk_b : 0;
        if (p.is_causal) kv_end = min(kv_end, max(0, mt + kBlockM + causal_off));
        nblk[t] = (kv_end + kBlockN - 1) / kBlockN;
    }
    const int n_blocks = max(nblk[0], nblk[1]);
    const int64_t o_row_base = (p.cu_q != nullptr) ? (int64_t)q_row0 : (int64_t)bidb * p.sq;
    float* lse_row = p.lse + ((int64_t)bidb * p.h + bidh) * p.sq;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + L::kOffQ;
    uint8_t* sKV = smem + L::kOffKV;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* bar_q = bars;                     // [2]
    uint64_t* bar_kv_full = bars + 2;           // [kStages]
    uint64_t* bar_kv_empty = bars + 2 + kStages;
    uint64_t* bar_s_full = bars + 2 + 2 * kStages;   // [2]
    uint64_t* bar_p_full = bar_s_full + 2;           // [2 tiles][2 halves]: P_t columns [0,64) / [64,128) written
    uint64_t* bar_o_full = bar_p_full + 4;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o_full + 2);

    if (warp == 8) {
        if (lane == 0) {
            mbar_init(&bar_q[0], 1); mbar_init(&bar_q[1], 1);
            for (int i = 0; i < kStages; ++i) { mbar_init(&bar_kv_full[i], 1); mbar_init(&bar_kv_empty[i], 1); }
            for (int t = 0; t < 2; ++t) {
                mbar_init(&bar_s_full[t], 1);
                mbar_init(&bar_p_full[2 * t], kBlockM);
                mbar_init(&bar_p_full[2 * t + 1], kBlockM);
                mbar_init(&bar_o_full[t], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
        tmem_relinquish();
    } else if (warp == 9 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wg == 2) {
        setmaxnreg_dec<72>();
        if (warp == 9) {
            // ===================== TMA producer =====================
            if (lane == 0 && n_blocks > 0) {
                auto load_kv = [&](const CUtensorMap* tm, int i, int j) {
                    const int slot = i % kStages;
                    mbar_wait(&bar_kv_empty[slot], ((i / kStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bar_kv_full[slot], L::kTile);
                    for (int s = 0; s < kSlabs; ++s)
                        tma_load_4d(sKV + slot * L::kTile + s * L::kSlab, tm, &bar_kv_full[slot], s * 64, bidh_k,
                                    k_row0 + j * kBlockN, tma_b);
                };
                auto load_q = [&](int t) {
                    mbar_arrive_expect_tx(&bar_q[t], L::kTile);
                    for (int s = 0; s < kSlabs; ++s)
                        tma_load_4d(sQ + t * L::kTile + s * L::kSlab, &tmQ, &bar_q[t], s * 64, bidh,
                                    q_row0 + m0 + t * kBlockM, tma_b);
                };
                if (nblk[0] > 0) load_q(0);
                load_kv(&tmK, 0, 0);
                if (nblk[1] > 0) load_q(1);
                load_kv(&tmV, 1, 0);
                for (int j = 1; j < n_blocks; ++j) {
                    load_kv(&tmK, 2 * j, j);
                    load_kv(&tmV, 2 * j + 1, j);
                }
            }
        } else if (warp == 8) {
            // ===================== MMA issuer =====================
            // The whole warp walks the (warp-uniform) schedule and waits on the barriers; one elected lane issues
            // the tcgen05 instructions.  Everything that feeds a descriptor is made provably warp-uniform
            // (__shfl_sync broadcast) so the compiler keeps it in uniform registers instead of emitting a
            // per-MMA "waterfall" loop — the issue rate of this warp bounds the whole kernel.
            const int nb0 = __shfl_sync(0xffffffffu, nblk[0], 0);
            const int nb1 = __shfl_sync(0xffffffffu, nblk[1], 0);
            const int nbmax = max(nb0, nb1);
            if (nbmax > 0) {
                const bool leader = elect_one();
                constexpr uint32_t idesc_s = make_idesc(kBf16, kBlockM, kBlockN, false, false);
                constexpr uint32_t idesc_pv = make_idesc(kBf16, kBlockM, D, false, true);
                const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
                const uint32_t q_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sQ), 16), 0);
                const uint32_t kv_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sKV), 16), 0);
                const uint32_t v_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sKV), L::kSlab), 0);
                constexpr uint32_t kTile16 = L::kTile >> 4;   // descriptor address units are 16 bytes
                auto issue_s = [&](int t, int j) {  // S_t = Q_t K_j^T
                    if (leader) {
                        const uint32_t qa = q_lo + t * kTile16;
                        const uint32_t ka = kv_lo + ((2 * j) % kStages) * kTile16;
#pragma unroll
                        for (int kk = 0; kk < D / 16; ++kk) {
                            const uint32_t off = ((kk >> 2) * L::kSlab + (kk & 3) * 32) >> 4;
                            umma_ss(tm + kTmemS0 + t * 128, desc_make(qa + off, kDescHiK), desc_make(ka + off, kDescHiK),
                                    idesc_s, kk > 0);
                        }
                        tc_commit(&bar_s_full[t]);
                    }
                };
                auto issue_pv = [&](int t, int j, int half) {  // O_t += P_t[:, 64 half : 64 half + 64] V_j[64 half ...]
                    if (leader) {
                        const uint32_t va = v_lo + ((2 * j + 1) % kStages) * kTile16;
#pragma unroll
                        for (int kk = half * 4; kk < half * 4 + 4; ++kk) {
                            umma_ts(tm + kTmemO0 + t * 128, tm + kTmemS0 + t * 128 + kk * 8,
                                    desc_make(va + kk * (2048 >> 4), kDescHiK), idesc_pv, (j > 0 || kk > 0));
                        }
                    }
                };
                auto wait_kv = [&](int i) { mbar_wait(&bar_kv_full[i % kStages], (i / kStages) & 1); };
                auto commit = [&](uint64_t* bar) { if (leader) tc_commit(bar); };

                wait_kv(0);
                tc_fence_after();
                if (nb0 > 0) { mbar_wait(&bar_q[0], 0); issue_s(0, 0); }
                if (nb1 > 0) { mbar_wait(&bar_q[1], 0); issue_s(1, 0); }
                commit(&bar_kv_empty[0]);
                for (int j = 0; j < nbmax; ++j) {
                    wait_kv(2 * j + 1);  // V_j
                    bool k_ready = false;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int nbt = t == 0 ? nb0 : nb1;
                        if (j < nbt) {
                            mbar_wait(&bar_p_full[2 * t], j & 1);
                            tc_fence_after();
                            if (lane == 0) FA_TRACE_EVENT(2, j, t);
                            issue_pv(t, j, 0);
                            mbar_wait(&bar_p_full[2 * t + 1], j & 1);
                            tc_fence_after();
                            issue_pv(t, j, 1);
                            if (j + 1 < nbt) {
                                if (!k_ready) { wait_kv(2 * j + 2); tc_fence_after(); k_ready = true; }
                                issue_s(t, j + 1);
                            } else {
                                commit(&bar_o_full[t]);
                            }
                            if (lane == 0) FA_TRACE_EVENT(2, j, 2 + t);
                        }
                    }
                    commit(&bar_kv_empty[(2 * j + 1) % kStages]);
                    if (j + 1 < nbmax) commit(&bar_kv_empty[(2 * j + 2) % kStages]);
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================== softmax warpgroups =====================
        setmaxnreg_inc<216>();
        const int t = wg;                     // query tile handled by this warpgroup
        const int r_in_tile = tid & 127;
        const int mt = m0 + t * kBlockM;
        const int row = mt + r_in_tile;
        const int n_t = nblk[t];
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_base + kTmemS0 + t * 128;
        const uint32_t tO = tmem_base + lane_base + kTmemO0 + t * 128;
        uint16_t* o_base = reinterpret_cast<uint16_t*>(p.o);
        constexpr int kChunksPerRow = D / 8;

        if (n_t == 0) {
            // no visible key for any row of this tile: O = 0, LSE = 0 (rows beyond seqlen_q are skipped)
            if (mt < sq_b) {
                for (int idx = r_in_tile; idx < kBlockM * kChunksPerRow; idx += kBlockM) {
                    const int r = idx / kChunksPerRow, ch = idx % kChunksPerRow;
                    if (mt + r < sq_b)
                        *(reinterpret_cast<uint4*>(o_base + ((o_row_base + mt + r) * p.h + bidh) * D) + ch) = make_uint4(0, 0, 0, 0);
                }
                if (row < sq_b) lse_row[row] = 0.f;
            }
        } else {
            int col_limit = sk_b - 1;
            if (p.is_causal) col_limit = min(col_limit, row + causal_off);
            float m_ref = -INFINITY, l_run = 0.f;
            const float c2 = p.scale_log2;

            for (int j = 0; j < n_t; ++j) {
                const int n0 = j * kBlockN;
                mbar_wait(&bar_s_full[t], j & 1);
                tc_fence_after();
                if (r_in_tile == 0) FA_TRACE_EVENT(t, j, 0);
                float s[kBlockN];
#pragma unroll
                for (int c = 0; c < kBlockN / 32; ++c)
                    tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
                tmem_wait_ld();
                if (r_in_tile == 0) FA_TRACE_EVENT(t, j, 1);

                const bool need_mask = (n0 + kBlockN > sk_b) || (p.is_causal && (n0 + kBlockN - 1 > mt + causal_off));
                if (need_mask) {
                    const int lim = col_limit - n0;
#pragma unroll
                    for (int c = 0; c < kBlockN; ++c)
                        if (c > lim) s[c] = -INFINITY;
                }
                softmax_step<D, kBf16, kEmu>(s, j == 0, c2, p.inv_scale_log2, m_ref, l_run, tS, tO, &bar_p_full[2 * t],
                                             &bar_p_full[2 * t + 1]);
                if (r_in_tile == 0) FA_TRACE_EVENT(t, j, 4);
            }

            // ---- epilogue for tile t ----
            mbar_wait(&bar_o_full[t], 0);
            tc_fence_after();
            const bool row_empty = (m_ref == -INFINITY) || !(l_run > 0.f);   // no visible key: O = 0, LSE = 0
            const float inv_l = row_empty ? 0.f : (1.f / l_run);
            uint8_t* sO = sQ + t * L::kTile;   // Q_t is dead: every S_t MMA retired before o_full[t]
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
                uint32_t o[32];
                tmem_ld32(tO + c * 32, o);
                tmem_wait_ld();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 v;
                    v.x = pack2<kBf16>(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
                    v.y = pack2<kBf16>(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
                    v.z = pack2<kBf16>(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
                    v.w = pack2<kBf16>(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
                    const int chunk = c * 4 + g;
                    *reinterpret_cast<uint4*>(sO + r_in_tile * (D * 2) + ((chunk ^ (r_in_tile & 7)) * 16)) = v;
                }
            }
            if (row < sq_b) lse_row[row] = row_empty ? 0.f : (m_ref * p.scale + logf(l_run));
            tc_fence_before();
            named_bar_sync(1 + t, kBlockM);
#pragma unroll 4
            for (int idx = r_in_tile; idx < kBlockM * kChunksPerRow; idx += kBlockM) {
                const int r = idx / kChunksPerRow, ch = idx % kChunksPerRow;
                if (mt + r < sq_b) {
                    const uint4 v = *reinterpret_cast<const uint4*>(sO + r * (D * 2) + ((ch ^ (r & 7)) * 16));
                    *(reinterpret_cast<uint4*>(o_base + ((o_row_base + mt + r) * p.h + bidh) * D) + ch) = v;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int D, bool kBf16>
static int launch_fwd_g(const fa_fwd_params* p, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                        const FwdParams& kp, cudaStream_t stream) {
    using L = FwdSmemG<D>;
    auto kern = flash_fwd_kernel_sm100_g<D, kBf16>;
    static bool attr_set = false;  // benign race: idempotent
    if (!attr_set) {
        FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes));
        attr_set = true;
    }
    dim3 grid((unsigned)((p->seqlen_q + kBlockM - 1) / kBlockM), (unsigned)p->h, (unsigned)p->b);
    kern<<<grid, 192, L::kBytes, stream>>>(tq, tk, tv, kp);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return FA_OK;
}

template <int D, bool kBf16, int kEmu>
static int launch_fwd_ws(const fa_fwd_params* p, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                         const FwdParams& kp, cudaStream_t stream) {
    using L = FwdSmem<D>;
    auto kern = flash_fwd_kernel_sm100<D, kBf16, kEmu>;
    static bool attr_set = false;
    if (!attr_set) {
        FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes));
        attr_set = true;
    }
    dim3 grid((unsigned)((p->seqlen_q + 2 * kBlockM - 1) / (2 * kBlockM)), (unsigned)p->h, (unsigned)p->b);
    kern<<<grid, 384, L::kBytes, stream>>>(tq, tk, tv, kp);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return FA_OK;
}

static int fwd_variant() {
    // FA_B200_FWD selects a kernel for A/B runs and debugging:
    //   (unset) / "p4" = 3: persistent kernel with four softmax warpgroups for head_dim 128 (flash_fwd_p4_sm100.cu),
    //                       the two-warpgroup persistent kernel for head_dim 64
    //   "ws"           = 0: two-warpgroup persistent kernel for every head_dim (flash_fwd_persist_sm100.cu)
    //   "np"           = 2: one CTA per work item (non-persistent) warp-specialised kernel
    //   "g"            = 1: single-tile bring-up kernel
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FA_B200_FWD");
        v = (e && e[0] == 'g') ? 1 : (e && e[0] == 'n') ? 2 : (e && e[0] == 'w') ? 0 : 3;
    }
    return v;
}

static int fwd_emu() {
    // FA_B200_EMU: share of the exponentials evaluated by a degree-3 polynomial on the FMA pipe instead of MUFU.EX2
    //   0 = none, 4 = 1 pair in 8 (p4 only), 1 = 2 in 8, 3 = 3 in 8 (p4 only), 2 = 4 in 8.
    // Default: 1 for the four-warpgroup kernel (measured +3 % over 0 and better than 4 / 3 / 2 at C2, C3 and C4, burst and
    // sustained), 0 for the two-warpgroup kernels (there 1 is +1.4 % in a burst and -0.5 % under the power cap).
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FA_B200_EMU");
        v = e ? atoi(e) : (fwd_variant() == 3 ? 1 : 0);
        if (v < 0 || v > 4) v = 0;
    }
    return v;
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
    kp.scale = 1.0f / sqrtf((float)p->d);
    kp.scale_log2 = kp.scale * 1.4426950408889634f;
    kp.inv_scale_log2 = 1.0f / kp.scale_log2;
    kp.trace = nullptr;
#ifdef FA_TRACE
    static long long* d_trace = nullptr;
    const bool do_trace = getenv("FA_B200_TRACE") != nullptr;
    if (do_trace) {
        if (!d_trace) cudaMalloc(&d_trace, 3 * 64 * 8 * sizeof(long long));
        cudaMemsetAsync(d_trace, 0, 3 * 64 * 8 * sizeof(long long), stream);
        kp.trace = d_trace;
    }
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
    if (fwd_variant() == 1) {
        if (p->d == 128) return bf16 ? launch_fwd_g<128, true>(p, tq, tk, tv, kp, stream) : launch_fwd_g<128, false>(p, tq, tk, tv, kp, stream);
        if (p->d == 64) return bf16 ? launch_fwd_g<64, true>(p, tq, tk, tv, kp, stream) : launch_fwd_g<64, false>(p, tq, tk, tv, kp, stream);
    }
#ifdef FA_TRACE
    if (do_trace && p->d == 128 && bf16) {
        int rc = fwd_variant() == 3 ? (fwd_emu() == 1 ? launch_fwd_p4<true, 1>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<true, 0>(p, tq, tk, tv, kp, stream))
               : fwd_variant() == 0 ? launch_fwd_persistent<128, true, 1>(p, tq, tk, tv, kp, stream)
                                    : launch_fwd_ws<128, true, 1>(p, tq, tk, tv, kp, stream);
        cudaStreamSynchronize(stream);
        static long long h[3 * 64 * 8];
        cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[0];
        printf("TRACE role j : events (cycles since first S-full of WG0)\n");
        for (int r = 0; r < 3; ++r)
            for (int j = 26; j < 40; ++j) {
                printf("TRACE %d %2d :", r, j);
                for (int e = 0; e < 8; ++e) printf(" %8lld", h[(r * 64 + j) * 8 + e] ? h[(r * 64 + j) * 8 + e] - t0 : -1LL);
                printf("\n");
            }
        fflush(stdout);
        return rc;
    }
#endif
    if (fwd_variant() == 3 && p->d == 128) {   // four softmax warpgroups (flash_fwd_p4_sm100.cu)
        switch (fwd_emu()) {
            case 0: return bf16 ? launch_fwd_p4<true, 0>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<false, 0>(p, tq, tk, tv, kp, stream);
            case 1: return bf16 ? launch_fwd_p4<true, 1>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<false, 1>(p, tq, tk, tv, kp, stream);
            case 3: return bf16 ? launch_fwd_p4<true, 3>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<false, 3>(p, tq, tk, tv, kp, stream);
            case 4: return bf16 ? launch_fwd_p4<true, 4>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<false, 4>(p, tq, tk, tv, kp, stream);
            default: return bf16 ? launch_fwd_p4<true, 2>(p, tq, tk, tv, kp, stream) : launch_fwd_p4<false, 2>(p, tq, tk, tv, kp, stream);
        }
    }
    if (fwd_variant() == 0 || fwd_variant() == 3) {   // persistent kernel (flash_fwd_persist_sm100.cu)
        if (p->d == 128) {
            switch (fwd_emu()) {
                case 0: return bf16 ? launch_fwd_persistent<128, true, 0>(p, tq, tk, tv, kp, stream) : launch_fwd_persistent<128, false, 0>(p, tq, tk, tv, kp, stream);
                case 1: return bf16 ? launch_fwd_persistent<128, true, 1>(p, tq, tk, tv, kp, stream) : launch_fwd_persistent<128, false, 1>(p, tq, tk, tv, kp, stream);
                default: return bf16 ? launch_fwd_persistent<128, true, 2>(p, tq, tk, tv, kp, stream) : launch_fwd_persistent<128, false, 2>(p, tq, tk, tv, kp, stream);
            }
        }
        if (p->d == 64) return bf16 ? launch_fwd_persistent<64, true, 0>(p, tq, tk, tv, kp, stream) : launch_fwd_persistent<64, false, 0>(p, tq, tk, tv, kp, stream);
    }
    if (p->d == 128) {
        switch (fwd_emu()) {
            case 0: return bf16 ? launch_fwd_ws<128, true, 0>(p, tq, tk, tv, kp, stream) : launch_fwd_ws<128, false, 0>(p, tq, tk, tv, kp, stream);
            case 1: return bf16 ? launch_fwd_ws<128, true, 1>(p, tq, tk, tv, kp, stream) : launch_fwd_ws<128, false, 1>(p, tq, tk, tv, kp, stream);
            default: return bf16 ? launch_fwd_ws<128, true, 2>(p, tq, tk, tv, kp, stream) : launch_fwd_ws<128, false, 2>(p, tq, tk, tv, kp, stream);
        }
    }
    if (p->d == 64) return bf16 ? launch_fwd_ws<64, true, 0>(p, tq, tk, tv, kp, stream) : launch_fwd_ws<64, false, 0>(p, tq, tk, tv, kp, stream);
    set_error("head_dim %lld not supported (64 or 128)", (long long)p->d);
    return FA_ERR_INVALID_ARG;
}

}  // namespace fa100
