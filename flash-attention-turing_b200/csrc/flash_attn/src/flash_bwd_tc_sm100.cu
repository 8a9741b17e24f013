// flash_bwd_tc_sm100.cu — tensor-core (tcgen05 / TMEM / TMA) attention backward for sm_100a.
//
// Same mathematical split as the reference (flash_bwd_kernel.h:28-838 dQ, :842-1676 dK/dV): two deterministic kernels,
// no atomics, P and dS recomputed from LSE and D = rowsum(dO*O):
//
//   flash_bwd_dq_kernel_sm100     one CTA per 128-row Q tile, loops over 128-row K/V tiles
//        S  = Q K^T, dP = dO V^T          (SS MMAs, fp32 in TMEM)
//        P  = exp2(S c2 - LSE c1),  dS = P * (dP - D)        (one thread per query row, two warpgroups split columns)
//        dQ += dS K                        (dS from TMEM as A, K consumed in place as an MN-major B)
//   flash_bwd_dk_dv_kernel_sm100  one CTA per 128-row K/V tile and KV head, loops over the GQA group's heads and over
//        64-row Q sub-tiles, everything transposed so that the K/V rows sit on the TMEM lanes:
//        S^T = K Q^T, dP^T = V dO^T        (M = 128 kv rows, N = 64 q rows)
//        P^T, dS^T elementwise             (LSE / D broadcast per column from shared memory)
//        dV += P^T dO,  dK += dS^T Q       (P^T/dS^T from TMEM as A; dO/Q in place as MN-major B)
//      The GQA group sum the reference does with h-expanded buffers + torch::sum_out (flash_api.cpp:265-312) is just
//      continued accumulation in TMEM here.
//
// Unlike the forward, nothing is aliased in TMEM (S, dP, dS/P and the accumulators have their own columns), so the
// MMAs of step j+1 are issued as soon as step j's S/dP have been read into registers and run underneath the
// elementwise work of step j.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "fa_common.h"
#include "flash_bwd_params.h"
#include "sm100_ptx.cuh"

#ifndef FA_FUSED_D64_DEFAULT
#define FA_FUSED_D64_DEFAULT 1   // head_dim 64 takes the fused backward too (+17-22 % over the two deterministic kernels, profiles/r02_run27.log); FA_B200_BWD_D64=det / fused override
#endif
#ifndef FA_FUSED_RED_COLS64
#define FA_FUSED_RED_COLS64 16  // the same at head_dim 64 (a red.global instruction only carries 16 lanes x 4 bytes there)
#endif
#ifndef FA_FUSED_STAGES64
#define FA_FUSED_STAGES64 6     // Q / dO ring depth at head_dim 64 (shared memory is plentiful there)
#endif
#ifndef FA_FUSED_EMU64
#define FA_FUSED_EMU64 0        // head_dim 64: column pairs out of every 8 whose 2^x goes through the FMA-pipe polynomial
#endif
#ifndef FA_FUSED_RED_COLS
#define FA_FUSED_RED_COLS 16   // query rows per step whose dQ leaves through red.global (the rest: one bulk reduction)
#endif

namespace fa100 {

constexpr int kBM = 128;
constexpr float kLog2e = 1.4426950408889634f;

struct SeqGeom {
    int q_row0, k_row0, sq_b, sk_b, tma_b;
};
FA_DEVICE SeqGeom seq_geom(const BwdParams& p, int bidb) {
    SeqGeom g;
    if (p.cu_q != nullptr) {
        g.q_row0 = p.cu_q[bidb]; g.sq_b = p.cu_q[bidb + 1] - g.q_row0;
        g.k_row0 = p.cu_k[bidb]; g.sk_b = p.cu_k[bidb + 1] - g.k_row0;
        // LSE / dsum rows and the fused dQ accumulator are sized by max_seqlen_q: never index past them
        g.sq_b = max(0, min(g.sq_b, p.sq)); g.sk_b = max(0, min(g.sk_b, p.sk));
        g.tma_b = 0;
    } else {
        g.q_row0 = 0; g.k_row0 = 0; g.sq_b = p.sq; g.sk_b = p.sk; g.tma_b = bidb;
    }
    return g;
}

// =================================================================================================
// dQ
// =================================================================================================
template <int D> struct DqSmem {
    static constexpr int kSlab = kBM * 128;
    static constexpr int kTile = kBM * D * 2;
    static constexpr int kKStages = 3;               // K_j is needed until dQ_j retires -> 3 deep
    static constexpr int kVStages = 2;               // V_j is released as soon as dP_j retires -> 2 deep
    static constexpr int kOffQ = 0;
    static constexpr int kOffDO = kTile;
    static constexpr int kOffK = 2 * kTile;
    static constexpr int kOffV = kOffK + kKStages * kTile;
    static constexpr int kOffBar = kOffV + kVStages * kTile;
    static constexpr int kBytes = kOffBar + 256 + 1024;   // 224 KB of tiles + barriers at D = 128
};
namespace dqt { constexpr uint32_t kS = 0, kDP = 128, kDQ = 256, kDS = 384; }

// kEmu (FA_B200_BWD_EMU=1, off by default): 2 of every 8 column pairs evaluate P = 2^x by polynomial on the FMA pipe
template <int D, bool kBf16, bool kEmu>
__global__ void __launch_bounds__(384, 1)
flash_bwd_dq_kernel_sm100(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                          const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                          const BwdParams p) {
    using L = DqSmem<D>;
    constexpr int kSlabs = D / 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = warp >> 2;
    const int mblk = p.is_causal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
    const int m0 = mblk * kBM;
    const int bidh = blockIdx.y, bidb = blockIdx.z, bidh_k = bidh / p.hratio;
    const SeqGeom sg = seq_geom(p, bidb);
    if (m0 >= sg.sq_b) return;
    const int off = sg.sk_b - sg.sq_b;
    int kv_end = sg.sk_b;
    if (p.is_causal) kv_end = min(sg.sk_b, max(0, m0 + kBM + off));
    const int nblk = (kv_end + kBM - 1) / kBM;
    const int64_t row_base = (p.cu_q != nullptr) ? (int64_t)sg.q_row0 : (int64_t)bidb * p.sq;
    uint16_t* dq_base = reinterpret_cast<uint16_t*>(p.dq);
    constexpr int kChunksPerRow = D / 8;

    if (nblk == 0) {  // no visible key: dQ = 0
        for (int idx = tid; idx < kBM * kChunksPerRow; idx += blockDim.x) {
            const int rr = idx / kChunksPerRow, ch = idx % kChunksPerRow;
            if (m0 + rr < sg.sq_b)
                *(reinterpret_cast<uint4*>(dq_base + ((row_base + m0 + rr) * p.h + bidh) * D) + ch) = make_uint4(0, 0, 0, 0);
        }
        return;
    }

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + L::kOffQ;
    uint8_t* sDO = smem + L::kOffDO;
    uint8_t* sK = smem + L::kOffK;
    uint8_t* sV = smem + L::kOffV;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* bar_q = bars;            // Q + dO landed
    uint64_t* bar_k_full = bars + 1;   // [3]
    uint64_t* bar_k_empty = bars + 4;  // [3]
    uint64_t* bar_v_full = bars + 7;   // [2]
    uint64_t* bar_v_empty = bars + 9;  // [2]
    uint64_t* bar_s_full = bars + 11;  // S and dP of step j complete
    uint64_t* bar_s_empty = bars + 12; // 256 threads have S/dP of step j in registers
    uint64_t* bar_ds_full = bars + 13; // 256 threads wrote dS_j
    uint64_t* bar_ds_empty = bars + 14;// dQ MMA of step j retired
    uint64_t* bar_dq_full = bars + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    if (warp == 8) {
        if (lane == 0) {
            mbar_init(bar_q, 1);
            for (int i = 0; i < L::kKStages; ++i) { mbar_init(&bar_k_full[i], 1); mbar_init(&bar_k_empty[i], 1); }
            for (int i = 0; i < L::kVStages; ++i) { mbar_init(&bar_v_full[i], 1); mbar_init(&bar_v_empty[i], 1); }
            mbar_init(bar_s_full, 1); mbar_init(bar_s_empty, 256);
            mbar_init(bar_ds_full, 256); mbar_init(bar_ds_empty, 1);
            mbar_init(bar_dq_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
        tmem_relinquish();
    } else if (warp == 9 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wg == 2) {
        setmaxnreg_dec<72>();
        if (warp == 9) {
            if (lane == 0) {
                mbar_arrive_expect_tx(bar_q, 2 * L::kTile);
                for (int s = 0; s < kSlabs; ++s) {
                    tma_load_4d(sQ + s * L::kSlab, &tmQ, bar_q, s * 64, bidh, sg.q_row0 + m0, sg.tma_b);
                    tma_load_4d(sDO + s * L::kSlab, &tmDO, bar_q, s * 64, bidh, sg.q_row0 + m0, sg.tma_b);
                }
                for (int j = 0; j < nblk; ++j) {
                    const int ks = j % L::kKStages, vs = j % L::kVStages;
                    mbar_wait(&bar_k_empty[ks], ((j / L::kKStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bar_k_full[ks], L::kTile);
                    for (int s = 0; s < kSlabs; ++s)
                        tma_load_4d(sK + ks * L::kTile + s * L::kSlab, &tmK, &bar_k_full[ks], s * 64, bidh_k, sg.k_row0 + j * kBM, sg.tma_b);
                    mbar_wait(&bar_v_empty[vs], ((j / L::kVStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bar_v_full[vs], L::kTile);
                    for (int s = 0; s < kSlabs; ++s)
                        tma_load_4d(sV + vs * L::kTile + s * L::kSlab, &tmV, &bar_v_full[vs], s * 64, bidh_k, sg.k_row0 + j * kBM, sg.tma_b);
                }
            }
        } else if (warp == 8) {
            const int nb = __shfl_sync(0xffffffffu, nblk, 0);
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc(kBf16, kBM, kBM, false, false);
            constexpr uint32_t idesc_dq = make_idesc(kBf16, kBM, D, false, true);
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t q_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sQ), 16), 0);
            const uint32_t do_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sDO), 16), 0);
            const uint32_t k_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sK), 16), 0);
            const uint32_t v_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sV), 16), 0);
            const uint32_t kmn_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sK), L::kSlab), 0);
            constexpr uint32_t kTile16 = L::kTile >> 4;
            auto issue_sdp = [&](int j) {
                const int sk_ = j % L::kKStages, sv_ = j % L::kVStages;
                mbar_wait(&bar_k_full[sk_], (j / L::kKStages) & 1);
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t o16 = ((kk >> 2) * L::kSlab + (kk & 3) * 32) >> 4;
                        umma_ss(tm + dqt::kS, desc_make(q_lo + o16, kDescHiK), desc_make(k_lo + sk_ * kTile16 + o16, kDescHiK),
                                idesc_s, kk > 0);
                    }
                }
                mbar_wait(&bar_v_full[sv_], (j / L::kVStages) & 1);
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t o16 = ((kk >> 2) * L::kSlab + (kk & 3) * 32) >> 4;
                        umma_ss(tm + dqt::kDP, desc_make(do_lo + o16, kDescHiK), desc_make(v_lo + sv_ * kTile16 + o16, kDescHiK),
                                idesc_s, kk > 0);
                    }
                    tc_commit(bar_s_full);
                    tc_commit(&bar_v_empty[sv_]);      // V_j is dead once dP_j has retired
                }
            };
            mbar_wait(bar_q, 0);
            issue_sdp(0);
            for (int j = 0; j < nb; ++j) {
                if (j + 1 < nb) {
                    mbar_wait(bar_s_empty, j & 1);
                    tc_fence_after();
                    if (lane == 0) FA_BTRACE(1, j, 0);
                    issue_sdp(j + 1);
                    if (lane == 0) FA_BTRACE(1, j, 1);
                }
                mbar_wait(bar_ds_full, j & 1);
                tc_fence_after();
                if (lane == 0) FA_BTRACE(1, j, 2);
                if (leader) {
                    const uint32_t ka = kmn_lo + (j % L::kKStages) * kTile16;
#pragma unroll
                    for (int kk = 0; kk < kBM / 16; ++kk)
                        umma_ts(tm + dqt::kDQ, tm + dqt::kDS + kk * 8, desc_make(ka + kk * (2048 >> 4), kDescHiK), idesc_dq,
                                (j > 0 || kk > 0));
                    tc_commit(bar_ds_empty);
                    tc_commit(&bar_k_empty[j % L::kKStages]);
                    if (j + 1 == nb) tc_commit(bar_dq_full);
                }
                if (lane == 0) FA_BTRACE(1, j, 3);
                __syncwarp();
            }
        }
    } else {
        // ============ elementwise warpgroups: thread (g, r) owns query row r and key columns [64 g, 64 g + 64) ============
        setmaxnreg_inc<216>();
        const int g = wg;
        const int r = ((warp & 3) << 5) | lane;
        const int row = m0 + r;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_base + dqt::kS + g * 64;
        const uint32_t tDP = tmem_base + lane_base + dqt::kDP + g * 64;
        const uint32_t tDS = tmem_base + lane_base + dqt::kDS + g * 32;
        const int64_t stat = ((int64_t)bidb * p.h + bidh) * p.sq + row;
        const bool row_ok = row < sg.sq_b;
        const float lse_r = row_ok ? p.lse[stat] : INFINITY;       // +inf => P = 0 for rows beyond the sequence
        const float d_r = row_ok ? p.dsum[stat] : 0.f;
        const float c2 = p.scale * kLog2e;
        const float neg = -lse_r * kLog2e;
        int col_limit = sg.sk_b - 1;
        if (p.is_causal) col_limit = min(col_limit, row + off);
        const float2 c2v = make_float2(c2, c2), negv = make_float2(neg, neg), ndv = make_float2(-d_r, -d_r);

        for (int j = 0; j < nblk; ++j) {
            const int n0 = j * kBM;
            mbar_wait(bar_s_full, j & 1);
            tc_fence_after();
            if (tid == 0) FA_BTRACE(0, j, 0);
            float s[64], dp[64];
            tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
            tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
            tmem_ld32(tDP, *reinterpret_cast<uint32_t(*)[32]>(&dp[0]));
            tmem_ld32(tDP + 32, *reinterpret_cast<uint32_t(*)[32]>(&dp[32]));
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(bar_s_empty);
            if (tid == 0) FA_BTRACE(0, j, 1);

            const bool need_mask = (n0 + kBM > sg.sk_b) || (p.is_causal && (n0 + kBM - 1 > m0 + off));
            if (need_mask) {
                const int lim = col_limit - n0 - g * 64;
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c > lim) s[c] = -INFINITY;
            }
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float2 x = __ffma2_rn(make_float2(s[2 * i], s[2 * i + 1]), c2v, negv);
                const float2 pr = (kEmu && (i & 3) == 0) ? exp2_poly_pair(x) : make_float2(fast_exp2(x.x), fast_exp2(x.y));
                const float2 ds = __fmul2_rn(pr, __fadd2_rn(make_float2(dp[2 * i], dp[2 * i + 1]), ndv));
                pk[i] = pack2<kBf16>(ds.x, ds.y);
            }
            if (tid == 0) FA_BTRACE(0, j, 2);
            if (j > 0) mbar_wait(bar_ds_empty, (j - 1) & 1);
            if (tid == 0) FA_BTRACE(0, j, 3);
            tmem_st32(tDS, pk);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar_ds_full);
            if (tid == 0) FA_BTRACE(0, j, 4);
        }

        // ---- epilogue: dQ * scale -> 16 bit -> smem (dead Q tile) -> coalesced stores ----
        mbar_wait(bar_dq_full, 0);
        tc_fence_after();
        constexpr int kHalfD = D / 2;
        const uint32_t tDQ = tmem_base + lane_base + dqt::kDQ + g * kHalfD;
        uint8_t* sO = sQ;
#pragma unroll
        for (int c = 0; c < kHalfD / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tDQ + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                uint4 v;
                v.x = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 0]) * p.scale, __uint_as_float(o[q4 * 8 + 1]) * p.scale);
                v.y = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 2]) * p.scale, __uint_as_float(o[q4 * 8 + 3]) * p.scale);
                v.z = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 4]) * p.scale, __uint_as_float(o[q4 * 8 + 5]) * p.scale);
                v.w = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 6]) * p.scale, __uint_as_float(o[q4 * 8 + 7]) * p.scale);
                const int chunk = g * (kHalfD / 8) + c * 4 + q4;
                *reinterpret_cast<uint4*>(sO + r * (D * 2) + ((chunk ^ (r & 7)) * 16)) = v;
            }
        }
        tc_fence_before();
        named_bar_sync(1, 256);
#pragma unroll 4
        for (int idx = tid; idx < kBM * kChunksPerRow; idx += 256) {
            const int rr = idx / kChunksPerRow, ch = idx % kChunksPerRow;
            if (m0 + rr < sg.sq_b) {
                const uint4 v = *reinterpret_cast<const uint4*>(sO + rr * (D * 2) + ((ch ^ (rr & 7)) * 16));
                *(reinterpret_cast<uint4*>(dq_base + ((row_base + m0 + rr) * p.h + bidh) * D) + ch) = v;
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

// =================================================================================================
// dK / dV
// =================================================================================================
constexpr int kSubQ = 64;   // query rows per step of the dK/dV kernel
template <int D> struct DkvSmem {
    static constexpr int kSlab = kBM * 128;            // 64-column slab of the 128-row K/V tiles
    static constexpr int kTile = kBM * D * 2;
    static constexpr int kSubSlab = kSubQ * 128;       // 64-column slab of a 64-row Q/dO sub-tile (8 KB)
    static constexpr int kSub = kSubQ * D * 2;
    static constexpr int kStages = 4;                  // each stage: Q sub-tile + dO sub-tile (TMA runs ~2 steps ahead)
    static constexpr int kOffK = 0;
    static constexpr int kOffV = kTile;
    static constexpr int kOffQdO = 2 * kTile;
    static constexpr int kOffStat = kOffQdO + kStages * 2 * kSub;   // float [2 buffers][2][64]: -LSE*log2e, D
    static constexpr int kOffBar = kOffStat + 2 * 2 * kSubQ * 4;
    static constexpr int kBytes = kOffBar + 256 + 1024;
};
namespace kvt { constexpr uint32_t kSt = 0, kDPt = 64, kPt = 128, kDSt = 160, kSt1 = 192, kDV = 256, kDK = 384; }   // S^T double-buffered: step st uses kSt (even) / kSt1 (odd)

template <int D, bool kBf16>
__global__ void __launch_bounds__(384, 1)
flash_bwd_dk_dv_kernel_sm100(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                             const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                             const BwdParams p) {
    using L = DkvSmem<D>;
    constexpr int kSlabs = D / 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = warp >> 2;
    const int n0 = blockIdx.x * kBM;
    const int bidh_k = blockIdx.y, bidb = blockIdx.z;
    const SeqGeom sg = seq_geom(p, bidb);
    if (n0 >= sg.sk_b) return;
    const int off = sg.sk_b - sg.sq_b;
    // query rows that can see any key of this tile: i >= n0 - off (causal), in 64-row sub-tiles
    const int i_first = p.is_causal ? max(0, n0 - off) : 0;
    const int it0 = i_first / kSubQ;
    const int nsub = (sg.sq_b + kSubQ - 1) / kSubQ;
    const int steps_per_head = max(0, nsub - it0);
    const int total = steps_per_head * p.hratio;
    const int64_t krow_base = (p.cu_q != nullptr) ? (int64_t)sg.k_row0 : (int64_t)bidb * p.sk;
    uint16_t* dk_base = reinterpret_cast<uint16_t*>(p.dk);
    uint16_t* dv_base = reinterpret_cast<uint16_t*>(p.dv);
    constexpr int kChunksPerRow = D / 8;

    if (total == 0) {  // no query row sees these keys: dK = dV = 0
        for (int idx = tid; idx < kBM * kChunksPerRow; idx += blockDim.x) {
            const int rr = idx / kChunksPerRow, ch = idx % kChunksPerRow;
            if (n0 + rr < sg.sk_b) {
                const int64_t o = ((krow_base + n0 + rr) * p.h_k + bidh_k) * D;
                *(reinterpret_cast<uint4*>(dk_base + o) + ch) = make_uint4(0, 0, 0, 0);
                *(reinterpret_cast<uint4*>(dv_base + o) + ch) = make_uint4(0, 0, 0, 0);
            }
        }
        return;
    }

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem + L::kOffK;
    uint8_t* sV = smem + L::kOffV;
    uint8_t* sQdO = smem + L::kOffQdO;
    float* sStat = reinterpret_cast<float*>(smem + L::kOffStat);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* bar_kv = bars;              // K + V landed
    uint64_t* bar_qdo_full = bars + 1;                   // [kStages]
    uint64_t* bar_qdo_empty = bars + 1 + L::kStages;     // [kStages]
    uint64_t* bar_s_full = bars + 1 + 2 * L::kStages;
    uint64_t* bar_s_empty = bar_s_full + 1;     // 256
    uint64_t* bar_p_full = bar_s_full + 2;      // 256
    uint64_t* bar_p_empty = bar_s_full + 3;
    uint64_t* bar_acc_full = bar_s_full + 4;
    uint64_t* bar_stat_full = bar_s_full + 5;   // [2] 64 arrivals (warps 10-11 published the column statistics)
    uint64_t* bar_stat_empty = bar_s_full + 7;  // [2] 256 arrivals (every elementwise thread has them in registers)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s_full + 9);

    if (warp == 8) {
        if (lane == 0) {
            mbar_init(bar_kv, 1);
            for (int i = 0; i < L::kStages; ++i) { mbar_init(&bar_qdo_full[i], 1); mbar_init(&bar_qdo_empty[i], 1); }
            mbar_init(bar_s_full, 1); mbar_init(bar_s_empty, 256);
            mbar_init(bar_p_full, 256); mbar_init(bar_p_empty, 1);
            mbar_init(bar_acc_full, 1);
            for (int i = 0; i < 2; ++i) { mbar_init(&bar_stat_full[i], kSubQ); mbar_init(&bar_stat_empty[i], 256); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
        tmem_relinquish();
    } else if (warp == 9 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wg == 2) {
        setmaxnreg_dec<72>();
        if (warp == 9) {
            if (lane == 0) {
                mbar_arrive_expect_tx(bar_kv, 2 * L::kTile);
                for (int s = 0; s < kSlabs; ++s) {
                    tma_load_4d(sK + s * L::kSlab, &tmK, bar_kv, s * 64, bidh_k, sg.k_row0 + n0, sg.tma_b);
                    tma_load_4d(sV + s * L::kSlab, &tmV, bar_kv, s * 64, bidh_k, sg.k_row0 + n0, sg.tma_b);
                }
                for (int st = 0; st < total; ++st) {
                    const int stage = st % L::kStages;
                    const int hq = bidh_k * p.hratio + st / steps_per_head;
                    const int it = it0 + st % steps_per_head;
                    mbar_wait(&bar_qdo_empty[stage], ((st / L::kStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bar_qdo_full[stage], 2 * L::kSub);
                    uint8_t* dst = sQdO + stage * 2 * L::kSub;
                    for (int s = 0; s < kSlabs; ++s) {
                        tma_load_4d(dst + s * L::kSubSlab, &tmQ, &bar_qdo_full[stage], s * 64, hq, sg.q_row0 + it * kSubQ, sg.tma_b);
                        tma_load_4d(dst + L::kSub + s * L::kSubSlab, &tmDO, &bar_qdo_full[stage], s * 64, hq,
                                    sg.q_row0 + it * kSubQ, sg.tma_b);
                    }
                }
            }
        } else if (warp == 8) {
            const int tot = __shfl_sync(0xffffffffu, total, 0);
            const bool leader = elect_one();
            constexpr uint32_t idesc_st = make_idesc(kBf16, kBM, kSubQ, false, false);   // M=128 kv, N=64 q
            constexpr uint32_t idesc_acc = make_idesc(kBf16, kBM, D, false, true);       // N = D, B MN-major
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t k_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sK), 16), 0);
            const uint32_t v_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sV), 16), 0);
            const uint32_t qk_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sQdO), 16), 0);             // K-major view
            const uint32_t qmn_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sQdO), L::kSubSlab), 0);   // MN-major view
            constexpr uint32_t kStage16 = (2 * L::kSub) >> 4, kSub16 = L::kSub >> 4;
            // S^T has two TMEM buffers, dP^T one.  S^T of step st+2 is issued right behind the accumulator MMAs of step st
            // (its buffer was released by bar_s_empty(st)), so the tensor pipe never waits for the elementwise warpgroups
            // to pick up step st+1 before it has more work: per step  dP^T(st+1) | dV,dK(st) | S^T(st+2).
            auto issue_s = [&](int st) {
                const int stage = st % L::kStages;
                mbar_wait(&bar_qdo_full[stage], (st / L::kStages) & 1);
                tc_fence_after();
                if (leader) {
                    const uint32_t qa = qk_lo + stage * kStage16;
                    const uint32_t ts = tm + ((st & 1) ? kvt::kSt1 : kvt::kSt);
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t oa = ((kk >> 2) * L::kSlab + (kk & 3) * 32) >> 4;
                        const uint32_t ob = ((kk >> 2) * L::kSubSlab + (kk & 3) * 32) >> 4;
                        umma_ss(ts, desc_make(k_lo + oa, kDescHiK), desc_make(qa + ob, kDescHiK), idesc_st, kk > 0);
                    }
                }
            };
            auto issue_dp = [&](int st) {   // the Q/dO stage of step st is already known to have landed (issue_s(st) waited)
                const int stage = st % L::kStages;
                if (leader) {
                    const uint32_t da = qk_lo + stage * kStage16 + kSub16;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t oa = ((kk >> 2) * L::kSlab + (kk & 3) * 32) >> 4;
                        const uint32_t ob = ((kk >> 2) * L::kSubSlab + (kk & 3) * 32) >> 4;
                        umma_ss(tm + kvt::kDPt, desc_make(v_lo + oa, kDescHiK), desc_make(da + ob, kDescHiK), idesc_st, kk > 0);
                    }
                    tc_commit(bar_s_full);      // S^T(st) was issued earlier on the same pipe: it has retired too
                }
            };
            mbar_wait(bar_kv, 0);
            issue_s(0);
            issue_dp(0);
            if (tot > 1) issue_s(1);
            for (int st = 0; st < tot; ++st) {
                if (st + 1 < tot) {
                    mbar_wait(bar_s_empty, st & 1);    // S^T(st), dP^T(st) are in registers: dP^T and S^T buffer (st & 1) are free
                    tc_fence_after();
                    if (lane == 0) FA_BTRACE(1, st, 0);
                    issue_dp(st + 1);
                    if (lane == 0) FA_BTRACE(1, st, 1);
                }
                mbar_wait(bar_p_full, st & 1);
                tc_fence_after();
                if (lane == 0) FA_BTRACE(1, st, 2);
                if (leader) {
                    const uint32_t qa = qmn_lo + (st % L::kStages) * kStage16, da = qa + kSub16;
#pragma unroll
                    for (int kk = 0; kk < kSubQ / 16; ++kk) {  // dV += P^T dO and dK += dS^T Q, interleaved
                        umma_ts(tm + kvt::kDV, tm + kvt::kPt + kk * 8, desc_make(da + kk * (2048 >> 4), kDescHiK), idesc_acc,
                                (st > 0 || kk > 0));
                        umma_ts(tm + kvt::kDK, tm + kvt::kDSt + kk * 8, desc_make(qa + kk * (2048 >> 4), kDescHiK), idesc_acc,
                                (st > 0 || kk > 0));
                    }
                    tc_commit(bar_p_empty);
                    tc_commit(&bar_qdo_empty[st % L::kStages]);
                    if (st + 1 == tot) tc_commit(bar_acc_full);
                }
                if (st + 2 < tot) issue_s(st + 2);
                if (lane == 0) FA_BTRACE(1, st, 3);
                __syncwarp();
            }
        } else {
            // ===================== warps 10-11: column statistics loader =====================
            // thread c of the 64 stages -LSE*log2e and -D of query row (sub-tile row c) for every step, two buffers ahead
            const int c = tid - 320;
            auto fetch = [&](int st, float& lse_raw, float& d_raw, bool& ok) {   // issue the two global loads, no use yet
                const int hq = bidh_k * p.hratio + st / steps_per_head;
                const int i = (it0 + st % steps_per_head) * kSubQ + c;
                ok = i < sg.sq_b;
                const int64_t o = ((int64_t)bidb * p.h + hq) * p.sq + (ok ? i : 0);
                lse_raw = p.lse[o];
                d_raw = p.dsum[o];
            };
            float lse_cur, d_cur, lse_nxt = 0.f, d_nxt = 0.f;
            bool ok_cur, ok_nxt = false;
            fetch(0, lse_cur, d_cur, ok_cur);
            for (int st = 0; st < total; ++st) {
                const int buf = st & 1;
                if (st + 1 < total) fetch(st + 1, lse_nxt, d_nxt, ok_nxt);   // one step ahead: latency hidden behind the wait
                mbar_wait(&bar_stat_empty[buf], ((st >> 1) & 1) ^ 1);
                float* dst = sStat + buf * 2 * kSubQ;
                dst[c] = ok_cur ? (-lse_cur * kLog2e) : -INFINITY;      // -inf => P = 0 for query rows beyond the sequence
                dst[kSubQ + c] = ok_cur ? -d_cur : 0.f;
                mbar_arrive(&bar_stat_full[buf]);
                lse_cur = lse_nxt; d_cur = d_nxt; ok_cur = ok_nxt;
            }
        }
    } else {
        // ===== elementwise warpgroups: thread (g, r) owns key row r and query columns [32 g, 32 g + 32) of the sub-tile =====
        setmaxnreg_inc<216>();
        const int g = wg;
        const int r = ((warp & 3) << 5) | lane;
        const int jg = n0 + r;                       // global key row of this thread
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tSt0 = tmem_base + lane_base + kvt::kSt + g * 32;
        const uint32_t tSt1 = tmem_base + lane_base + kvt::kSt1 + g * 32;
        const uint32_t tDPt = tmem_base + lane_base + kvt::kDPt + g * 32;
        const uint32_t tPt = tmem_base + lane_base + kvt::kPt + g * 16;
        const uint32_t tDSt = tmem_base + lane_base + kvt::kDSt + g * 16;
        const float c2 = p.scale * kLog2e;
        const float2 c2v = make_float2(c2, c2);

        for (int st = 0; st < total; ++st) {
            const int it = it0 + st % steps_per_head;
            const int buf = st & 1;
            mbar_wait(bar_s_full, st & 1);
            tc_fence_after();
            if (tid == 0) FA_BTRACE(0, st, 0);
            float s[32], dp[32];
            tmem_ld32((st & 1) ? tSt1 : tSt0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
            tmem_ld32(tDPt, *reinterpret_cast<uint32_t(*)[32]>(&dp[0]));
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(bar_s_empty);                // lets the MMA warp issue S^T/dP^T of the next step right away
            // per-column statistics (-LSE*log2e, -D) of this sub-tile, staged by warps 10-11 (broadcast LDS.128).
            // They are read AFTER the arrive on purpose: all the arithmetic below depends on them, which keeps the
            // scheduler from sinking the TMEM loads / the arrive underneath the exponentials (measured: that delayed the
            // next step's MMAs by ~700 cycles).
            if (tid == 0) FA_BTRACE(0, st, 4);
            mbar_wait(&bar_stat_full[buf], (st >> 1) & 1);
            if (tid == 0) FA_BTRACE(0, st, 5);
            const uint32_t stat = smem_u32(sStat) + (buf * 2 * kSubQ + g * 32) * 4;
            float4 nl4[8], dd4[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                nl4[i] = lds128(stat + i * 16);
                dd4[i] = lds128(stat + kSubQ * 4 + i * 16);
            }

            // causal: key jg visible to query i iff jg <= i + off  <=>  column c >= jg - off - it*64 - 32 g
            const bool need_mask = p.is_causal && (it * kSubQ < n0 + kBM - 1 - off);
            const int cmin = need_mask ? (jg - off - it * kSubQ - g * 32) : 0;
            uint32_t pkp[16], pkd[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 nl = (i & 1) ? make_float2(nl4[i >> 1].z, nl4[i >> 1].w) : make_float2(nl4[i >> 1].x, nl4[i >> 1].y);
                const float2 dd = (i & 1) ? make_float2(dd4[i >> 1].z, dd4[i >> 1].w) : make_float2(dd4[i >> 1].x, dd4[i >> 1].y);
                const float2 x = __ffma2_rn(make_float2(s[2 * i], s[2 * i + 1]), c2v, nl);
                float2 pr = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                if (need_mask) {
                    if (2 * i < cmin) pr.x = 0.f;
                    if (2 * i + 1 < cmin) pr.y = 0.f;
                }
                const float2 ds = __fmul2_rn(pr, __fadd2_rn(make_float2(dp[2 * i], dp[2 * i + 1]), dd));
                pkp[i] = pack2<kBf16>(pr.x, pr.y);
                pkd[i] = pack2<kBf16>(ds.x, ds.y);
            }
            if (tid == 0) FA_BTRACE(0, st, 1);
            mbar_arrive(&bar_stat_empty[buf]);       // statistics consumed
            if (st > 0) mbar_wait(bar_p_empty, (st - 1) & 1);
            if (tid == 0) FA_BTRACE(0, st, 2);
            tmem_st16(tPt, pkp);
            tmem_st16(tDSt, pkd);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar_p_full);
            if (tid == 0) FA_BTRACE(0, st, 3);
        }

        // ---- epilogue: dV, dK * scale -> 16 bit -> smem (dead V / K tiles) -> coalesced stores ----
        mbar_wait(bar_acc_full, 0);
        tc_fence_after();
        constexpr int kHalfD = D / 2;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const uint32_t tA = tmem_base + lane_base + (which == 0 ? kvt::kDV : kvt::kDK) + g * kHalfD;
            uint8_t* sO = which == 0 ? sV : sK;
            const float mul = which == 0 ? 1.f : p.scale;
#pragma unroll
            for (int c = 0; c < kHalfD / 32; ++c) {
                uint32_t o[32];
                tmem_ld32(tA + c * 32, o);
                tmem_wait_ld();
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 v;
                    v.x = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 0]) * mul, __uint_as_float(o[q4 * 8 + 1]) * mul);
                    v.y = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 2]) * mul, __uint_as_float(o[q4 * 8 + 3]) * mul);
                    v.z = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 4]) * mul, __uint_as_float(o[q4 * 8 + 5]) * mul);
                    v.w = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 6]) * mul, __uint_as_float(o[q4 * 8 + 7]) * mul);
                    const int chunk = g * (kHalfD / 8) + c * 4 + q4;
                    *reinterpret_cast<uint4*>(sO + r * (D * 2) + ((chunk ^ (r & 7)) * 16)) = v;
                }
            }
        }
        tc_fence_before();
        named_bar_sync(1, 256);
#pragma unroll 2
        for (int idx = tid; idx < kBM * kChunksPerRow; idx += 256) {
            const int rr = idx / kChunksPerRow, ch = idx % kChunksPerRow;
            if (n0 + rr < sg.sk_b) {
                const int64_t o = ((krow_base + n0 + rr) * p.h_k + bidh_k) * D;
                const int so = rr * (D * 2) + ((ch ^ (rr & 7)) * 16);
                *(reinterpret_cast<uint4*>(dv_base + o) + ch) = *reinterpret_cast<const uint4*>(sV + so);
                *(reinterpret_cast<uint4*>(dk_base + o) + ch) = *reinterpret_cast<const uint4*>(sK + so);
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


// =================================================================================================
// fused backward (default for head_dim 128 when the caller provides the workspace; FA_B200_BWD=det selects the two
// deterministic kernels above): the dK/dV kernel also produces dQ, 5 GEMMs per tile pair instead of 7
// =================================================================================================
// Same K/V-stationary, transposed formulation as flash_bwd_dk_dv_kernel_sm100 (with a single S^T buffer), plus
//     dQ^T[d x 64 q] = K^T dS^T      (A = the resident K tile read MN-major, B = dS^T from shared memory, MN-major)
// per step, drained by the elementwise warpgroups into an fp32 accumulator [rows, h, d] in global memory with
// red.global.add.f32 (coalesced: one q row x 32 consecutive d per warp instruction); flash_bwd_dq_kernel_sm100_convert
// scales and rounds it to dQ afterwards.  Unlike everything else in this library the result depends on the order in
// which the K/V tiles' contributions land (fp32 reductions in L2): dQ is not bit-reproducible run to run (dK and dV
// are).  Measured +11 % (C2) to +17 % (C3, C4) over the two-kernel path; the kernel is bound by the L2's fp32 reduction
// rate (~2 KB/clk for the whole chip, however the reductions are issued: scalar red 2.2, 512-byte bulk 1.3, 16 KB bulk 2.1).
template <int D> struct FzSmem {
    static constexpr int kSlab = kBM * 128;
    static constexpr int kTile = kBM * D * 2;
    static constexpr int kSubSlab = kSubQ * 128;
    static constexpr int kSub = kSubQ * D * 2;
    static constexpr int kStages = D == 64 ? FA_FUSED_STAGES64 : 3;                 // D = 128: one less than the plain kernel (pays for the dQ staging tile)
    static constexpr int kOffK = 0;
    static constexpr int kOffV = kTile;
    static constexpr int kOffQdO = 2 * kTile;
    static constexpr int kOffStat = kOffQdO + kStages * 2 * kSub;   // float [2 buffers][2][64]
    static constexpr int kOffDS = kOffStat + 1024;                  // 2 x 16 KB, 1024-byte aligned
    static constexpr int kOffDQ = kOffDS + 2 * kBM * kSubQ * 2;     // fp32 [64 q rows][128 d]: source of the bulk reductions
    static constexpr int kOffBar = kOffDQ + kSubQ * D * 4;
    static constexpr int kBytes = kOffBar + 256 + 1024;
};
static_assert(FzSmem<128>::kBytes <= 232448 && FzSmem<64>::kBytes <= 232448, "fused backward: shared memory budget");
namespace fzt { constexpr uint32_t kSt = 0, kDPt = 64, kPt = 128, kDSt = 160, kDQt = 192, kDV = 256, kDK = 384; }

// kAcc16 (FA_B200_BWD_ACC=16, experiment): the partial dQ tiles are reduced in fp16 instead of fp32 (half the L2 reduction
// bytes; the running sums are rounded to 11 bits at every one of the sk/128 additions)
template <int D, bool kBf16, bool kAcc16>
__global__ void __launch_bounds__(512, 1)
flash_bwd_dk_dv_kernel_sm100_fused(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                             const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                             const BwdParams p) {
    using L = FzSmem<D>;
    constexpr int kSlabs = D / 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = warp >> 2;
    const int n0 = blockIdx.x * kBM;
    const int bidh_k = blockIdx.y, bidb = blockIdx.z;
    const SeqGeom sg = seq_geom(p, bidb);
    if (n0 >= sg.sk_b) return;
    const int off = sg.sk_b - sg.sq_b;
    // query rows that can see any key of this tile: i >= n0 - off (causal), in 64-row sub-tiles
    const int i_first = p.is_causal ? max(0, n0 - off) : 0;
    const int it0 = i_first / kSubQ;
    const int nsub = (sg.sq_b + kSubQ - 1) / kSubQ;
    const int steps_per_head = max(0, nsub - it0);
    const int total = steps_per_head * p.hratio;
    // The CTAs of one head (different K/V tiles) would all reduce into the SAME dQ rows at the same time if they walked
    // the query sub-tiles in the same order — same-address reductions serialise in L2.  Each CTA therefore starts its walk
    // at a different sub-tile (the order is irrelevant: the contributions commute).
    const int rot = steps_per_head > 0 ? (int)((blockIdx.x * 2) % steps_per_head) : 0;
    auto q_sub = [&](int st) { int ls = st % steps_per_head + rot; if (ls >= steps_per_head) ls -= steps_per_head; return it0 + ls; };
    const int64_t krow_base = (p.cu_q != nullptr) ? (int64_t)sg.k_row0 : (int64_t)bidb * p.sk;
    uint16_t* dk_base = reinterpret_cast<uint16_t*>(p.dk);
    uint16_t* dv_base = reinterpret_cast<uint16_t*>(p.dv);
    constexpr int kChunksPerRow = D / 8;

    if (total == 0) {  // no query row sees these keys: dK = dV = 0
        for (int idx = tid; idx < kBM * kChunksPerRow; idx += blockDim.x) {
            const int rr = idx / kChunksPerRow, ch = idx % kChunksPerRow;
            if (n0 + rr < sg.sk_b) {
                const int64_t o = ((krow_base + n0 + rr) * p.h_k + bidh_k) * D;
                *(reinterpret_cast<uint4*>(dk_base + o) + ch) = make_uint4(0, 0, 0, 0);
                *(reinterpret_cast<uint4*>(dv_base + o) + ch) = make_uint4(0, 0, 0, 0);
            }
        }
        return;
    }

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem + L::kOffK;
    uint8_t* sV = smem + L::kOffV;
    uint8_t* sQdO = smem + L::kOffQdO;
    float* sStat = reinterpret_cast<float*>(smem + L::kOffStat);
    uint8_t* sDS = smem + L::kOffDS;      // dS^T (16 bit) [2 buffers][128 kv rows][64 q = 128 B], 128B-swizzled: B operand of dQ^T
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* bar_kv = bars;              // K + V landed
    uint64_t* bar_qdo_full = bars + 1;                   // [kStages]
    uint64_t* bar_qdo_empty = bars + 1 + L::kStages;     // [kStages]
    uint64_t* bar_s_full = bars + 1 + 2 * L::kStages;
    uint64_t* bar_s_empty = bar_s_full + 1;     // 256
    uint64_t* bar_p_full = bar_s_full + 2;      // 256
    uint64_t* bar_p_empty = bar_s_full + 3;
    uint64_t* bar_acc_full = bar_s_full + 4;
    uint64_t* bar_stat_full = bar_s_full + 5;   // [2] 64 arrivals (warps 10-11 published the column statistics)
    uint64_t* bar_stat_empty = bar_s_full + 7;  // [2] 256 arrivals (every elementwise thread has them in registers)
    uint64_t* bar_dq_full = bar_s_full + 9;     // dQ^T of a step complete in TMEM
    uint64_t* bar_dq_empty = bar_s_full + 10;   // 256 threads have it in registers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s_full + 11);

    if (warp == 8) {
        if (lane == 0) {
#ifdef FA_HANG_GUARD
            g_hang_info[6] = smem_u32(bars); g_hang_info[7] = (uint32_t)total;
#endif
            mbar_init(bar_kv, 1);
            for (int i = 0; i < L::kStages; ++i) { mbar_init(&bar_qdo_full[i], 1); mbar_init(&bar_qdo_empty[i], 1); }
            mbar_init(bar_s_full, 1); mbar_init(bar_s_empty, 256);
            mbar_init(bar_p_full, 256); mbar_init(bar_p_empty, 1);
            mbar_init(bar_acc_full, 1);
            mbar_init(bar_dq_full, 1); mbar_init(bar_dq_empty, 128);
            for (int i = 0; i < 2; ++i) { mbar_init(&bar_stat_full[i], kSubQ); mbar_init(&bar_stat_empty[i], 256); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
        tmem_relinquish();
    } else if (warp == 9 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wg == 3) {
        // ===================== warpgroup 3: dQ^T drain (TMEM -> shared staging tile -> one bulk reduction per step) =====================
        // dQ^T of step s sits in TMEM as [d = lane][64 q columns]; thread r = d writes row c, element r of the staging tile
        // (32 lanes -> 128 contiguous bytes), then ONE thread issues a single cp.reduce.async.bulk of the whole tile: the
        // accumulator is head-major ([b][h][sq_pad][d]) so the 64 q rows of a step are contiguous in global memory.
        // (Tried first, see DESIGN.md: per-thread red.global.add.f32, per-row 512-byte bulk reductions, and draining from
        //  the elementwise warpgroups — all slower.)
        // head_dim 64: dQ^T is an M = 64 accumulator — row d sits on lane 16 (d / 16) ... of TMEM, i.e. on lanes 0-15 of
        // every warp's 32-lane quarter; the upper half of each drain warp holds nothing and only keeps the barriers company.
        setmaxnreg_dec<40>();
        const int tw = tid & 127;
        constexpr bool kM64 = (D == 64);
        const bool dvalid = !kM64 || lane < 16;
        const int r = kM64 ? (((warp & 3) << 4) | (lane & 15)) : (((warp & 3) << 5) | lane);
        const uint32_t tDQt = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + fzt::kDQt;
        const uint32_t sDQr = smem_u32(smem + L::kOffDQ) + r * (kAcc16 ? 2 : 4);
        // The bulk-reduction engine of an SM moves ~15 B/clk: 32 KB of fp32 per step would take longer than the step's MMAs.
        // The first kRedCols query rows of every step therefore leave through per-thread red.global.add.f32 straight from
        // registers (a different path: LSU -> L2), the rest through the staging tile and ONE bulk reduction.
        constexpr int kRedCols = kAcc16 ? 0 : (kM64 ? FA_FUSED_RED_COLS64 : FA_FUSED_RED_COLS);
        constexpr int kBulkRows = kSubQ - kRedCols;
        for (int s = 0; s < total; ++s) {
            const int hq = bidh_k * p.hratio + s / steps_per_head;
            const int64_t eoff = (((int64_t)bidb * p.h + hq) * p.sq_pad + (int64_t)q_sub(s) * kSubQ) * D;
            mbar_wait(bar_dq_full, s & 1);
            tc_fence_after();
#pragma unroll
            for (int qt = 0; qt < 4; ++qt) {                 // 16 columns at a time: this warpgroup lives on 40 registers
                uint32_t v[16];
                tmem_ld16(tDQt + qt * 16, v);
                tmem_wait_ld();
                if (qt == 3) {
                    tc_fence_before();
                    mbar_arrive(bar_dq_empty);               // dQ^T(s) is in registers: the next step's MMA may overwrite it
                }
                if (qt * 16 < kRedCols) {
                    float* dst = p.dqacc + eoff + (qt * 16) * D + r;
                    if (dvalid) {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + c * D), "f"(__uint_as_float(v[c])) : "memory");
                    }
                } else {
                    if (qt * 16 == kRedCols) {
                        if (tw == 0) tma_store_wait_read<0>();   // the previous step's bulk reduction has read the staging tile
                        named_bar_sync(2, 128);
                    }
                    if (dvalid) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            if constexpr (kAcc16) {
                                uint16_t hv;
                                asm("cvt.rn.f16.f32 %0, %1;" : "=h"(hv) : "f"(__uint_as_float(v[c])));
                                asm volatile("st.shared.b16 [%0], %1;" ::"r"(sDQr + (qt * 16 - kRedCols + c) * (D * 2)), "h"(hv) : "memory");
                            } else {
                                asm volatile("st.shared.b32 [%0], %1;" ::"r"(sDQr + (qt * 16 - kRedCols + c) * (D * 4)), "r"(v[c]) : "memory");
                            }
                        }
                    }
                }
            }
            fence_proxy_async_smem();
            named_bar_sync(2, 128);
            if (tw == 0) {
                const uint32_t sDQ = smem_u32(smem + L::kOffDQ);
                if constexpr (kAcc16) {
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.f16 [%0], [%1], %2;"
                                 ::"l"(reinterpret_cast<uint16_t*>(p.dqacc) + eoff), "r"(sDQ), "n"(kSubQ * D * 2) : "memory");
                } else {
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                 ::"l"(p.dqacc + eoff + kRedCols * D), "r"(sDQ), "n"(kBulkRows * D * 4) : "memory");
                }
                tma_store_commit();
            }
        }
        if (tw == 0) tma_store_wait<0>();            // all bulk reductions have completed before the CTA retires
    } else if (wg == 2) {
        setmaxnreg_dec<72>();
        if (warp == 9) {
            if (lane == 0) {
                mbar_arrive_expect_tx(bar_kv, 2 * L::kTile);
                for (int s = 0; s < kSlabs; ++s) {
                    tma_load_4d(sK + s * L::kSlab, &tmK, bar_kv, s * 64, bidh_k, sg.k_row0 + n0, sg.tma_b);
                    tma_load_4d(sV + s * L::kSlab, &tmV, bar_kv, s * 64, bidh_k, sg.k_row0 + n0, sg.tma_b);
                }
                for (int st = 0; st < total; ++st) {
                    const int stage = st % L::kStages;
                    const int hq = bidh_k * p.hratio + st / steps_per_head;
                    const int it = q_sub(st);
                    mbar_wait(&bar_qdo_empty[stage], ((st / L::kStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bar_qdo_full[stage], 2 * L::kSub);
                    uint8_t* dst = sQdO + stage * 2 * L::kSub;
                    for (int s = 0; s < kSlabs; ++s) {
                        tma_load_4d(dst + s * L::kSubSlab, &tmQ, &bar_qdo_full[stage], s * 64, hq, sg.q_row0 + it * kSubQ, sg.tma_b);
                        tma_load_4d(dst + L::kSub + s * L::kSubSlab, &tmDO, &bar_qdo_full[stage], s * 64, hq,
                                    sg.q_row0 + it * kSubQ, sg.tma_b);
                    }
                }
            }
        } else if (warp == 8) {
            const int tot = __shfl_sync(0xffffffffu, total, 0);
            const bool leader = elect_one();
            constexpr uint32_t idesc_st = make_idesc(kBf16, kBM, kSubQ, false, false);   // M=128 kv, N=64 q
            constexpr uint32_t idesc_acc = make_idesc(kBf16, kBM, D, false, true);       // N = D, B MN-major
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t k_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sK), 16), 0);
            const uint32_t v_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sV), 16), 0);
            const uint32_t qk_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sQdO), 16), 0);             // K-major view
            const uint32_t qmn_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sQdO), L::kSubSlab), 0);   // MN-major view
            constexpr uint32_t kStage16 = (2 * L::kSub) >> 4, kSub16 = L::kSub >> 4;
            constexpr uint32_t idesc_dq = make_idesc(kBf16, D, kSubQ, true, true);       // dQ^T: M = d, N = 64 q, A and B MN-major
            const uint32_t kmn_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sK), L::kSlab), 0);    // K^T as an MN-major A operand
            const uint32_t ds_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sDS), 16), 0);
            auto issue_sdp = [&](int st) {
                const int stage = st % L::kStages;
                mbar_wait(&bar_qdo_full[stage], (st / L::kStages) & 1);
                tc_fence_after();
                if (leader) {
                    const uint32_t qa = qk_lo + stage * kStage16, da = qa + kSub16;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t oa = ((kk >> 2) * L::kSlab + (kk & 3) * 32) >> 4;
                        const uint32_t ob = ((kk >> 2) * L::kSubSlab + (kk & 3) * 32) >> 4;
                        umma_ss(tm + fzt::kSt, desc_make(k_lo + oa, kDescHiK), desc_make(qa + ob, kDescHiK), idesc_st, kk > 0);
                        umma_ss(tm + fzt::kDPt, desc_make(v_lo + oa, kDescHiK), desc_make(da + ob, kDescHiK), idesc_st, kk > 0);
                    }
                    tc_commit(bar_s_full);
                }
            };
            mbar_wait(bar_kv, 0);
            issue_sdp(0);
            for (int st = 0; st < tot; ++st) {
                if (st + 1 < tot) {
                    mbar_wait(bar_s_empty, st & 1);
                    tc_fence_after();
                    if (lane == 0) FA_BTRACE(1, st, 0);
                    issue_sdp(st + 1);
                    if (lane == 0) FA_BTRACE(1, st, 1);
                }
                mbar_wait(bar_p_full, st & 1);
                tc_fence_after();
                if (lane == 0) FA_BTRACE(1, st, 2);
                if (leader) {
                    const uint32_t qa = qmn_lo + (st % L::kStages) * kStage16, da = qa + kSub16;
#pragma unroll
                    for (int kk = 0; kk < kSubQ / 16; ++kk) {  // dV += P^T dO and dK += dS^T Q, interleaved
                        umma_ts(tm + fzt::kDV, tm + fzt::kPt + kk * 8, desc_make(da + kk * (2048 >> 4), kDescHiK), idesc_acc,
                                (st > 0 || kk > 0));
                        umma_ts(tm + fzt::kDK, tm + fzt::kDSt + kk * 8, desc_make(qa + kk * (2048 >> 4), kDescHiK), idesc_acc,
                                (st > 0 || kk > 0));
                    }
                    tc_commit(bar_p_empty);                          // P^T / dS^T in TMEM are free once these retire
                    tc_commit(&bar_qdo_empty[st % L::kStages]);
                }
                // dQ^T(st) [d x 64 q] = K^T dS^T(st): its TMEM tile must have been drained by the elementwise warps
                if (st > 0) { mbar_wait(bar_dq_empty, (st - 1) & 1); tc_fence_after(); }
                if (leader) {
                    const uint32_t dsa = ds_lo + (st & 1) * (uint32_t)((kBM * kSubQ * 2) >> 4);
#pragma unroll
                    for (int kk = 0; kk < kBM / 16; ++kk)
                        umma_ss(tm + fzt::kDQt, desc_make(kmn_lo + kk * (2048 >> 4), kDescHiK), desc_make(dsa + kk * (2048 >> 4), kDescHiK),
                                idesc_dq, kk > 0);
                    tc_commit(bar_dq_full);
                    if (st + 1 == tot) tc_commit(bar_acc_full);
                }
                if (lane == 0) FA_BTRACE(1, st, 3);
                __syncwarp();
            }
        } else {
            // ===================== warps 10-11: column statistics loader =====================
            // thread c of the 64 stages -LSE*log2e and -D of query row (sub-tile row c) for every step, two buffers ahead
            const int c = tid - 320;
            auto fetch = [&](int st, float& lse_raw, float& d_raw, bool& ok) {   // issue the two global loads, no use yet
                const int hq = bidh_k * p.hratio + st / steps_per_head;
                const int i = q_sub(st) * kSubQ + c;
                ok = i < sg.sq_b;
                const int64_t o = ((int64_t)bidb * p.h + hq) * p.sq + (ok ? i : 0);
                lse_raw = p.lse[o];
                d_raw = p.dsum[o];
            };
            float lse_cur, d_cur, lse_nxt = 0.f, d_nxt = 0.f;
            bool ok_cur, ok_nxt = false;
            fetch(0, lse_cur, d_cur, ok_cur);
            for (int st = 0; st < total; ++st) {
                const int buf = st & 1;
                if (st + 1 < total) fetch(st + 1, lse_nxt, d_nxt, ok_nxt);   // one step ahead: latency hidden behind the wait
                mbar_wait(&bar_stat_empty[buf], ((st >> 1) & 1) ^ 1);
                float* dst = sStat + buf * 2 * kSubQ;
                dst[c] = ok_cur ? (-lse_cur * kLog2e) : -INFINITY;      // -inf => P = 0 for query rows beyond the sequence
                dst[kSubQ + c] = ok_cur ? -d_cur : 0.f;
                mbar_arrive(&bar_stat_full[buf]);
                lse_cur = lse_nxt; d_cur = d_nxt; ok_cur = ok_nxt;
            }
        }
    } else {
        // ===== elementwise warpgroups: thread (g, r) owns key row r and query columns [32 g, 32 g + 32) of the sub-tile =====
        setmaxnreg_inc<200>();
        const int g = wg;
        const int r = ((warp & 3) << 5) | lane;
        const int jg = n0 + r;                       // global key row of this thread
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tSt = tmem_base + lane_base + fzt::kSt + g * 32;
        const uint32_t tDPt = tmem_base + lane_base + fzt::kDPt + g * 32;
        const uint32_t tPt = tmem_base + lane_base + fzt::kPt + g * 16;
        const uint32_t tDSt = tmem_base + lane_base + fzt::kDSt + g * 16;
        const float c2 = p.scale * kLog2e;
        const float2 c2v = make_float2(c2, c2);

        for (int st = 0; st < total; ++st) {
            const int it = q_sub(st);
            const int buf = st & 1;
            mbar_wait(bar_s_full, st & 1);
            tc_fence_after();
            if (tid == 0) FA_BTRACE(0, st, 0);
            float s[32], dp[32];
            tmem_ld32(tSt, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
            tmem_ld32(tDPt, *reinterpret_cast<uint32_t(*)[32]>(&dp[0]));
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(bar_s_empty);                // lets the MMA warp issue S^T/dP^T of the next step right away
            // per-column statistics (-LSE*log2e, -D) of this sub-tile, staged by warps 10-11 (broadcast LDS.128).
            // They are read AFTER the arrive on purpose: all the arithmetic below depends on them, which keeps the
            // scheduler from sinking the TMEM loads / the arrive underneath the exponentials (measured: that delayed the
            // next step's MMAs by ~700 cycles).
            if (tid == 0) FA_BTRACE(0, st, 4);
            mbar_wait(&bar_stat_full[buf], (st >> 1) & 1);
            if (tid == 0) FA_BTRACE(0, st, 5);
            const uint32_t stat = smem_u32(sStat) + (buf * 2 * kSubQ + g * 32) * 4;
            float4 nl4[8], dd4[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                nl4[i] = lds128(stat + i * 16);
                dd4[i] = lds128(stat + kSubQ * 4 + i * 16);
            }

            // causal: key jg visible to query i iff jg <= i + off  <=>  column c >= jg - off - it*64 - 32 g
            const bool need_mask = p.is_causal && (it * kSubQ < n0 + kBM - 1 - off);
            const int cmin = need_mask ? (jg - off - it * kSubQ - g * 32) : 0;
            uint32_t pkp[16], pkd[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 nl = (i & 1) ? make_float2(nl4[i >> 1].z, nl4[i >> 1].w) : make_float2(nl4[i >> 1].x, nl4[i >> 1].y);
                const float2 dd = (i & 1) ? make_float2(dd4[i >> 1].z, dd4[i >> 1].w) : make_float2(dd4[i >> 1].x, dd4[i >> 1].y);
                const float2 x = __ffma2_rn(make_float2(s[2 * i], s[2 * i + 1]), c2v, nl);
                float2 pr = (D == 64 && (i & 7) < FA_FUSED_EMU64) ? exp2_poly_pair(x) : make_float2(fast_exp2(x.x), fast_exp2(x.y));
                if (need_mask) {
                    if (2 * i < cmin) pr.x = 0.f;
                    if (2 * i + 1 < cmin) pr.y = 0.f;
                }
                const float2 ds = __fmul2_rn(pr, __fadd2_rn(make_float2(dp[2 * i], dp[2 * i + 1]), dd));
                pkp[i] = pack2<kBf16>(pr.x, pr.y);
                pkd[i] = pack2<kBf16>(ds.x, ds.y);
            }
            if (jg >= sg.sk_b) {     // key rows beyond the sequence (varlen: the next sequence's rows) must not reach dQ
#pragma unroll
                for (int i = 0; i < 16; ++i) { pkp[i] = 0u; pkd[i] = 0u; }
            }
            if (tid == 0) FA_BTRACE(0, st, 1);
            mbar_arrive(&bar_stat_empty[buf]);       // statistics consumed
            if (st > 0) mbar_wait(bar_p_empty, (st - 1) & 1);
            if (tid == 0) FA_BTRACE(0, st, 2);
            tmem_st16(tPt, pkp);
            tmem_st16(tDSt, pkd);
            {   // dS^T also goes to shared memory (buffer st & 1) as the MN-major B operand of dQ^T: row r, 16-byte chunks 4g..4g+3
                const uint32_t row_addr = smem_u32(sDS) + (st & 1) * (kBM * kSubQ * 2) + r * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    sts128u(row_addr + (((g * 4 + c) ^ (r & 7)) << 4), make_uint4(pkd[4 * c], pkd[4 * c + 1], pkd[4 * c + 2], pkd[4 * c + 3]));
            }
            fence_proxy_async_smem();
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar_p_full);
            if (tid == 0) FA_BTRACE(0, st, 3);
        }

        // ---- epilogue: dV, dK * scale -> 16 bit -> smem (dead V / K tiles) -> coalesced stores ----
        mbar_wait(bar_acc_full, 0);
        tc_fence_after();
        constexpr int kHalfD = D / 2;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const uint32_t tA = tmem_base + lane_base + (which == 0 ? fzt::kDV : fzt::kDK) + g * kHalfD;
            uint8_t* sO = which == 0 ? sV : sK;
            const float mul = which == 0 ? 1.f : p.scale;
#pragma unroll
            for (int c = 0; c < kHalfD / 32; ++c) {
                uint32_t o[32];
                tmem_ld32(tA + c * 32, o);
                tmem_wait_ld();
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 v;
                    v.x = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 0]) * mul, __uint_as_float(o[q4 * 8 + 1]) * mul);
                    v.y = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 2]) * mul, __uint_as_float(o[q4 * 8 + 3]) * mul);
                    v.z = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 4]) * mul, __uint_as_float(o[q4 * 8 + 5]) * mul);
                    v.w = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 6]) * mul, __uint_as_float(o[q4 * 8 + 7]) * mul);
                    const int chunk = g * (kHalfD / 8) + c * 4 + q4;
                    *reinterpret_cast<uint4*>(sO + r * (D * 2) + ((chunk ^ (r & 7)) * 16)) = v;
                }
            }
        }
        tc_fence_before();
        named_bar_sync(1, 256);
#pragma unroll 2
        for (int idx = tid; idx < kBM * kChunksPerRow; idx += 256) {
            const int rr = idx / kChunksPerRow, ch = idx % kChunksPerRow;
            if (n0 + rr < sg.sk_b) {
                const int64_t o = ((krow_base + n0 + rr) * p.h_k + bidh_k) * D;
                const int so = rr * (D * 2) + ((ch ^ (rr & 7)) * 16);
                *(reinterpret_cast<uint4*>(dv_base + o) + ch) = *reinterpret_cast<const uint4*>(sV + so);
                *(reinterpret_cast<uint4*>(dk_base + o) + ch) = *reinterpret_cast<const uint4*>(sK + so);
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


// dQ[b, i, h, :] = round16(scale * acc[b, h, i, :]): D / 8 lanes x 8 elements per row, 2048 / D rows per block, HBM-bound
template <int D, bool kBf16, bool kAcc16>
__global__ void __launch_bounds__(256)
flash_bwd_dq_kernel_sm100_convert(const BwdParams p) {
    constexpr int kLanes = D / 8;
    const int bidb = blockIdx.z, bidh = blockIdx.y;
    const int i = blockIdx.x * (256 / kLanes) + (threadIdx.x / kLanes), ch = threadIdx.x % kLanes;
    int q_row0 = 0, sq_b = p.sq;
    if (p.cu_q) { q_row0 = p.cu_q[bidb]; sq_b = min(p.cu_q[bidb + 1] - q_row0, p.sq); }
    if (i >= sq_b) return;
    const int64_t row_base = p.cu_q ? (int64_t)q_row0 : (int64_t)bidb * p.sq;
    const int64_t eoff = (((int64_t)bidb * p.h + bidh) * p.sq_pad + i) * D + 8 * ch;
    float4 a, b;
    if constexpr (kAcc16) {
        const uint4 w = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.dqacc) + eoff);
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&w.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
        const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&w.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&w.w));
        a = make_float4(f0.x, f0.y, f1.x, f1.y); b = make_float4(f2.x, f2.y, f3.x, f3.y);
    } else {
        const float4* src = reinterpret_cast<const float4*>(p.dqacc + eoff);
        a = src[0]; b = src[1];
    }
    uint4 o;
    o.x = pack2<kBf16>(a.x * p.scale, a.y * p.scale); o.y = pack2<kBf16>(a.z * p.scale, a.w * p.scale);
    o.z = pack2<kBf16>(b.x * p.scale, b.y * p.scale); o.w = pack2<kBf16>(b.z * p.scale, b.w * p.scale);
    reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dq) + ((row_base + i) * p.h + bidh) * D)[ch] = o;
}

// =================================================================================================
// host
// =================================================================================================
template <int D, bool kBf16>
static int launch_tc(const BwdParams& kp, const CUtensorMap& tq128, const CUtensorMap& tdo128, const CUtensorMap& tq64,
                     const CUtensorMap& tdo64, const CUtensorMap& tk, const CUtensorMap& tv, cudaStream_t stream) {
    // FA_B200_BWD_EMU=1: polynomial exponentials for 2 of 8 column pairs in the dQ kernel.  Measured: the elementwise phase
    // drops from 1300 to 960 cycles per step, but the kernel is then bound by its MMA issue loop (1770 cycles per step
    // either way; 2.89 vs 2.85 ms for the whole backward at C2) -> off by default, MUFU.EX2 everywhere.
    static int dq_emu = -1;
    if (dq_emu < 0) { const char* e = getenv("FA_B200_BWD_EMU"); dq_emu = e ? (atoi(e) != 0) : 0; }
    auto kdq = dq_emu ? flash_bwd_dq_kernel_sm100<D, kBf16, true> : flash_bwd_dq_kernel_sm100<D, kBf16, false>;
    auto kkv = flash_bwd_dk_dv_kernel_sm100<D, kBf16>;
    const DeviceInfo* di = nullptr;
    int rc = current_device_info(&di);
    if (rc != FA_OK) return rc;
    static std::atomic<unsigned long long> m_dq1{0}, m_dq0{0}, m_kv{0};   // per device, see ensure_dynamic_smem
    if ((rc = ensure_dynamic_smem(flash_bwd_dq_kernel_sm100<D, kBf16, true>, DqSmem<D>::kBytes, m_dq1, di->ordinal)) != FA_OK) return rc;
    if ((rc = ensure_dynamic_smem(flash_bwd_dq_kernel_sm100<D, kBf16, false>, DqSmem<D>::kBytes, m_dq0, di->ordinal)) != FA_OK) return rc;
    if ((rc = ensure_dynamic_smem(kkv, DkvSmem<D>::kBytes, m_kv, di->ordinal)) != FA_OK) return rc;
    if (kp.sq > 0) {
        dim3 g((kp.sq + kBM - 1) / kBM, kp.h, kp.b);
        kdq<<<g, 384, DqSmem<D>::kBytes, stream>>>(tq128, tdo128, tk, tv, kp);
        FA_CUDA_CHECK(cudaGetLastError());
        count_launch();
    }
    if (kp.sk > 0) {
        dim3 g((kp.sk + kBM - 1) / kBM, kp.h_k, kp.b);
        kkv<<<g, 384, DkvSmem<D>::kBytes, stream>>>(tq64, tdo64, tk, tv, kp);
        FA_CUDA_CHECK(cudaGetLastError());
        count_launch();
    }
    return FA_OK;
}

bool bwd_fused_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FA_B200_BWD"); v = (!e || e[0] == 'f') ? 1 : 0; }   // default; "det" / "rows" select the others
    return v == 1;
}
// head_dim 64 (FA_B200_BWD_D64=det selects the two deterministic kernels instead): the same fused kernel with an M = 64 dQ^T
static bool bwd_fused_d64_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FA_B200_BWD_D64"); v = e ? (e[0] == 'f' ? 1 : 0) : FA_FUSED_D64_DEFAULT; }
    return v == 1;
}
static bool bwd_fused_for(long long d) { return bwd_fused_enabled() && (d == 128 || (d == 64 && bwd_fused_d64_enabled())); }
long long bwd_fused_workspace_bytes(long long b, long long sq_max, long long h, long long d) {
    // fp32 accumulator [b][h][sq_pad][d], sq_pad = sq_max rounded up to the 64-row step of the kernel
    return bwd_fused_for(d) ? b * h * ((sq_max + 63) / 64 * 64) * d * 4 : 0;
}

static bool bwd_acc16() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FA_B200_BWD_ACC"); v = (e && atoi(e) == 16) ? 1 : 0; }
    return v == 1;
}

template <int D, bool kBf16, bool kAcc16>
static int launch_fused(const BwdParams& kp, const CUtensorMap& tq64, const CUtensorMap& tdo64, const CUtensorMap& tk,
                        const CUtensorMap& tv, cudaStream_t stream) {
    auto kern = flash_bwd_dk_dv_kernel_sm100_fused<D, kBf16, kAcc16>;
    const DeviceInfo* di = nullptr;
    int rc = current_device_info(&di);
    if (rc != FA_OK) return rc;
    static std::atomic<unsigned long long> attr_mask{0};
    if ((rc = ensure_dynamic_smem(kern, FzSmem<D>::kBytes, attr_mask, di->ordinal)) != FA_OK) return rc;
    FA_CUDA_CHECK(cudaMemsetAsync(kp.dqacc, 0, (size_t)bwd_fused_workspace_bytes(kp.b, kp.sq, kp.h, D) / (kAcc16 ? 2 : 1), stream));
    dim3 g((kp.sk + kBM - 1) / kBM, kp.h_k, kp.b);
    kern<<<g, 512, FzSmem<D>::kBytes, stream>>>(tq64, tdo64, tk, tv, kp);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    constexpr int kRowsPerBlock = 256 / (D / 8);
    dim3 gc((kp.sq + kRowsPerBlock - 1) / kRowsPerBlock, kp.h, kp.b);
    flash_bwd_dq_kernel_sm100_convert<D, kBf16, kAcc16><<<gc, 256, 0, stream>>>(kp);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return FA_OK;
}

int launch_bwd_tc_sm100(const BwdParams& kp, bool bf16, cudaStream_t stream) {
    const bool varlen = kp.cu_q != nullptr;
    const uint64_t D = (uint64_t)kp.d;
    const uint64_t rows_q = varlen ? (uint64_t)kp.total_q : (uint64_t)kp.sq;
    const uint64_t rows_k = varlen ? (uint64_t)kp.total_k : (uint64_t)kp.sk;
    const uint64_t nb = varlen ? 1 : (uint64_t)kp.b;
    if (rows_q == 0 || rows_k == 0) return -1;   // degenerate: the caller zero-fills the gradients
    CUtensorMap tq128, tdo128, tq64, tdo64, tk, tv;
    const uint32_t box128[4] = {64, 1, 128, 1}, box64[4] = {64, 1, 64, 1};
    {
        const uint64_t dims[4] = {D, (uint64_t)kp.h, rows_q, nb};
        const uint64_t str[3] = {D * 2, (uint64_t)kp.h * D * 2, rows_q * (uint64_t)kp.h * D * 2};
        int rc;
        if ((rc = encode_tmap_4d(&tq128, kp.q, bf16, dims, str, box128)) != FA_OK) return rc;
        if ((rc = encode_tmap_4d(&tdo128, kp.dout, bf16, dims, str, box128)) != FA_OK) return rc;
        if ((rc = encode_tmap_4d(&tq64, kp.q, bf16, dims, str, box64)) != FA_OK) return rc;
        if ((rc = encode_tmap_4d(&tdo64, kp.dout, bf16, dims, str, box64)) != FA_OK) return rc;
    }
    {
        const uint64_t dims[4] = {D, (uint64_t)kp.h_k, rows_k, nb};
        const uint64_t str[3] = {D * 2, (uint64_t)kp.h_k * D * 2, rows_k * (uint64_t)kp.h_k * D * 2};
        int rc;
        if ((rc = encode_tmap_4d(&tk, kp.k, bf16, dims, str, box128)) != FA_OK) return rc;
        if ((rc = encode_tmap_4d(&tv, kp.v, bf16, dims, str, box128)) != FA_OK) return rc;
    }
    if (kp.d == 128 && bwd_fused_for(128) && kp.dqacc != nullptr)
        return bwd_acc16() ? (bf16 ? launch_fused<128, true, true>(kp, tq64, tdo64, tk, tv, stream) : launch_fused<128, false, true>(kp, tq64, tdo64, tk, tv, stream))
                           : (bf16 ? launch_fused<128, true, false>(kp, tq64, tdo64, tk, tv, stream) : launch_fused<128, false, false>(kp, tq64, tdo64, tk, tv, stream));
    if (kp.d == 64 && bwd_fused_for(64) && kp.dqacc != nullptr)
        return bf16 ? launch_fused<64, true, false>(kp, tq64, tdo64, tk, tv, stream) : launch_fused<64, false, false>(kp, tq64, tdo64, tk, tv, stream);
    if (kp.d == 128) return bf16 ? launch_tc<128, true>(kp, tq128, tdo128, tq64, tdo64, tk, tv, stream)
                                 : launch_tc<128, false>(kp, tq128, tdo128, tq64, tdo64, tk, tv, stream);
    if (kp.d == 64) return bf16 ? launch_tc<64, true>(kp, tq128, tdo128, tq64, tdo64, tk, tv, stream)
                                : launch_tc<64, false>(kp, tq128, tdo128, tq64, tdo64, tk, tv, stream);
    return -1;
}

}  // namespace fa100

#ifdef FA_HANG_GUARD
extern "C" int fa_b200_hang_read(unsigned int* out8) {   // diagnosis builds only; not part of include/fa_b200.h
    unsigned int flag = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&flag, fa100::g_hang_flag, sizeof(flag));
    cudaMemcpyFromSymbol(out8, fa100::g_hang_info, 8 * sizeof(unsigned int));
    return (int)flag;
}
#endif
