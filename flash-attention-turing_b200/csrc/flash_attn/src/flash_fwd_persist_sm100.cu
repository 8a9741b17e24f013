// flash_fwd_persist_sm100.cu — persistent variant of the warp-specialised forward (flash_fwd_kernel_sm100).
//
// Same tile math, roles and TMEM plan as flash_fwd_sm100.cu (two 128-row query tiles per work item, softmax
// warpgroups 0/1, MMA warp 8, TMA warp 9), but one CTA per SM stays resident and walks a static list of work items.
// What that buys (measured on the one-CTA-per-work-item kernel: ~18 % of the SM time was outside the steady-state
// loop — CTA launch, barrier/TMEM set-up, first TMA round trip, pipeline fill, epilogue, CTA drain):
//   * barrier init, TMEM allocation and descriptor prefetch happen once per SM instead of once per tile pair;
//   * the TMA warp runs ahead into the next work item (its Q tiles and first K/V tiles land while the current item's
//     last P V and epilogue are still in flight);
//   * the MMA warp issues the next item's S = Q K^T while the softmax warpgroups write out the previous O;
//   * the epilogue goes TMEM -> registers -> global directly (each thread owns one 256-byte output row), so the Q
//     tiles in shared memory are free for the next item as soon as their last S MMA has retired.
//
// Work order: (batch, head) pairs are taken in groups whose K/V fit comfortably in L2; inside a group the items are
// ordered by decreasing cost (causal: later row blocks first) with the heads interleaved, and dealt round-robin to the
// CTAs.  Every CTA therefore sees a sawtooth of costs that averages out (no atomics, deterministic), while the CTAs
// that run concurrently work on the same few heads.
#include <stdlib.h>

#include "flash_fwd_common.cuh"

namespace fa100 {

template <int D, bool kBf16, int kEmu>
__global__ void __launch_bounds__(384, 1)
flash_fwd_kernel_sm100_persistent(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                                  const FwdParams p, const TileSched ts) {
    using L = FwdSmemP<D>;
    constexpr int kSlabs = D / 64;
    constexpr int kStages = L::kKvStages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int wg = warp >> 2;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + L::kOffQ;
    uint8_t* sKV = smem + L::kOffKV;
    uint8_t* sStage = smem + L::kOffStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBarP);
    uint64_t* bar_q_full = bars;                      // [2]  Q_t landed
    uint64_t* bar_q_empty = bars + 2;                 // [2]  last S_t MMA of the item retired
    uint64_t* bar_kv_full = bars + 4;                 // [kStages]
    uint64_t* bar_kv_empty = bars + 4 + kStages;
    uint64_t* bar_s_full = bars + 4 + 2 * kStages;    // [2]
    uint64_t* bar_p_full = bar_s_full + 2;            // [2 tiles][2 halves]
    uint64_t* bar_o_full = bar_p_full + 4;            // [2]  last P V of the item retired
    uint64_t* bar_o_empty = bar_o_full + 2;           // [2]  epilogue has O_t in registers (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o_empty + 2);
    int* stage_lock = reinterpret_cast<int*>(tmem_slot + 1);   // the two epilogues share one staging tile

    if (warp == 8) {
        if (lane == 0) {
            *stage_lock = 0;
            for (int t = 0; t < 2; ++t) {
                mbar_init(&bar_q_full[t], 1); mbar_init(&bar_q_empty[t], 1);
                mbar_init(&bar_s_full[t], 1);
                mbar_init(&bar_p_full[2 * t], kBlockM); mbar_init(&bar_p_full[2 * t + 1], kBlockM);
                mbar_init(&bar_o_full[t], 1); mbar_init(&bar_o_empty[t], kBlockM);
            }
            for (int i = 0; i < kStages; ++i) { mbar_init(&bar_kv_full[i], 1); mbar_init(&bar_kv_empty[i], 1); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
        tmem_relinquish();
    } else if (warp == 9 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wg == 2) {
        setmaxnreg_dec<72>();
        if (warp == 9) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                int kv_i = 0;            // running K/V ring index
                int nq[2] = {0, 0};      // Q_t loads so far
                for (int n = blockIdx.x; n < ts.total; n += gridDim.x) {
                    const WorkItem w = decode_item(ts, n, p.h, p.is_causal != 0);
                    const ItemGeom g = item_geom(p, w);
                    if (g.skip || g.n_blocks == 0) continue;
                    const int bidh_k = w.bidh / p.hratio;
                    auto load_kv = [&](const CUtensorMap* tm, int j) {
                        const int slot = kv_i % kStages;
                        mbar_wait(&bar_kv_empty[slot], ((kv_i / kStages) & 1) ^ 1);
                        mbar_arrive_expect_tx(&bar_kv_full[slot], L::kTile);
                        for (int s = 0; s < kSlabs; ++s)
                            tma_load_4d(sKV + slot * L::kTile + s * L::kSlab, tm, &bar_kv_full[slot], s * 64, bidh_k,
                                        g.k_row0 + j * kBlockN, g.tma_b);
                        ++kv_i;
                    };
                    auto load_q = [&](int t) {
                        if (g.nblk[t] == 0) return;
                        mbar_wait(&bar_q_empty[t], (nq[t] & 1) ^ 1);
                        mbar_arrive_expect_tx(&bar_q_full[t], L::kTile);
                        for (int s = 0; s < kSlabs; ++s)
                            tma_load_4d(sQ + t * L::kTile + s * L::kSlab, &tmQ, &bar_q_full[t], s * 64, w.bidh,
                                        g.q_row0 + g.m0 + t * kBlockM, g.tma_b);
                        ++nq[t];
                    };
                    load_q(0);
                    load_kv(&tmK, 0);
                    load_q(1);
                    load_kv(&tmV, 0);
                    for (int j = 1; j < g.n_blocks; ++j) {
                        load_kv(&tmK, j);
                        load_kv(&tmV, j);
                    }
                }
            }
        } else if (warp == 8) {
            // ===================== MMA issuer (warp-uniform walk, one elected lane issues) =====================
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc(kBf16, kBlockM, kBlockN, false, false);
            constexpr uint32_t idesc_pv = make_idesc(kBf16, kBlockM, D, false, true);
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t q_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sQ), 16), 0);
            const uint32_t kv_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sKV), 16), 0);
            const uint32_t v_lo = __shfl_sync(0xffffffffu, desc_lo(smem_u32(sKV), L::kSlab), 0);
            constexpr uint32_t kTile16 = L::kTile >> 4;
            int kv_i = 0;                 // ring index of K_0 of the current item
            int it[2] = {0, 0};           // S_t / P_t steps so far (barrier parities)
            int nitem[2] = {0, 0};        // items finished per tile slot (Q / O barrier parities)
            for (int n = blockIdx.x; n < ts.total; n += gridDim.x) {
                const WorkItem w = decode_item(ts, n, p.h, p.is_causal != 0);
                const ItemGeom g = item_geom(p, w);
                const int nb0 = __shfl_sync(0xffffffffu, g.nblk[0], 0);
                const int nb1 = __shfl_sync(0xffffffffu, g.nblk[1], 0);
                const int nbmax = max(nb0, nb1);
                if (__shfl_sync(0xffffffffu, (int)g.skip, 0) || nbmax == 0) continue;
                auto kv_slot = [&](int i) { return (kv_i + i) % kStages; };
                auto wait_kv = [&](int i) { mbar_wait(&bar_kv_full[kv_slot(i)], (((kv_i + i) / kStages) & 1)); };
                auto commit = [&](uint64_t* bar) { if (leader) tc_commit(bar); };
                auto issue_s = [&](int t, int j, int nbt) {  // S_t = Q_t K_j^T
                    if (leader) {
                        const uint32_t qa = q_lo + t * kTile16;
                        const uint32_t ka = kv_lo + kv_slot(2 * j) * kTile16;
#pragma unroll
                        for (int kk = 0; kk < D / 16; ++kk) {
                            const uint32_t off = ((kk >> 2) * L::kSlab + (kk & 3) * 32) >> 4;
                            umma_ss(tm + kTmemS0 + t * 128, desc_make(qa + off, kDescHiK), desc_make(ka + off, kDescHiK),
                                    idesc_s, kk > 0);
                        }
                        tc_commit(&bar_s_full[t]);
                        if (j + 1 == nbt) tc_commit(&bar_q_empty[t]);   // Q_t may be overwritten by the next item
                    }
                };
                auto issue_pv = [&](int t, int j, int half) {  // O_t += P_t[:, half] V_j[half]
                    if (leader) {
                        const uint32_t va = v_lo + kv_slot(2 * j + 1) * kTile16;
#pragma unroll
                        for (int kk = half * 4; kk < half * 4 + 4; ++kk)
                            umma_ts(tm + kTmemO0 + t * 128, tm + kTmemS0 + t * 128 + kk * 8,
                                    desc_make(va + kk * (2048 >> 4), kDescHiK), idesc_pv, (j > 0 || kk > 0));
                    }
                };

                wait_kv(0);
                tc_fence_after();
                if (nb0 > 0) { mbar_wait(&bar_q_full[0], nitem[0] & 1); issue_s(0, 0, nb0); }
                if (nb1 > 0) { mbar_wait(&bar_q_full[1], nitem[1] & 1); issue_s(1, 0, nb1); }
                commit(&bar_kv_empty[kv_slot(0)]);
                for (int j = 0; j < nbmax; ++j) {
                    wait_kv(2 * j + 1);  // V_j
                    bool k_ready = false;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int nbt = t == 0 ? nb0 : nb1;
                        if (j < nbt) {
                            if (j == 0) {   // O_t of the previous item must have been read out by its epilogue
                                mbar_wait(&bar_o_empty[t], (nitem[t] & 1) ^ 1);
                            }
                            mbar_wait(&bar_p_full[2 * t], (it[t] + j) & 1);
                            tc_fence_after();
                            if (lane == 0) FA_TRACE_EVENT(2, it[0] + j, t);
                            issue_pv(t, j, 0);
                            mbar_wait(&bar_p_full[2 * t + 1], (it[t] + j) & 1);
                            tc_fence_after();
                            issue_pv(t, j, 1);
                            if (j + 1 < nbt) {
                                if (!k_ready) { wait_kv(2 * j + 2); tc_fence_after(); k_ready = true; }
                                issue_s(t, j + 1, nbt);
                            } else {
                                commit(&bar_o_full[t]);
                            }
                            if (lane == 0) FA_TRACE_EVENT(2, it[0] + j, 2 + t);
                        }
                    }
                    commit(&bar_kv_empty[kv_slot(2 * j + 1)]);
                    if (j + 1 < nbmax) commit(&bar_kv_empty[kv_slot(2 * j + 2)]);
                    __syncwarp();
                }
                kv_i += 2 * nbmax;
                it[0] += nb0; it[1] += nb1;
                nitem[0] += (nb0 > 0); nitem[1] += (nb1 > 0);
            }
        }
    } else {
        // ===================== softmax warpgroups: warpgroup t owns tile slot t, one thread per query row =====================
        setmaxnreg_inc<216>();
        const int t = wg;
        const int r_in_tile = tid & 127;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_base + kTmemS0 + t * 128;
        const uint32_t tO = tmem_base + lane_base + kTmemO0 + t * 128;
        uint16_t* o_base = reinterpret_cast<uint16_t*>(p.o);
        const float c2 = p.scale_log2;
        int its = 0;       // S_t steps so far
        int nitem = 0;     // items with keys finished by this slot

        for (int n = blockIdx.x; n < ts.total; n += gridDim.x) {
            const WorkItem w = decode_item(ts, n, p.h, p.is_causal != 0);
            const ItemGeom g = item_geom(p, w);
            if (g.skip) continue;
            const int mt = g.m0 + t * kBlockM;
            if (mt >= g.sq_b) continue;                      // this slot has no rows in this item (nblk[t] == 0 too)
            const int row = mt + r_in_tile;
            const int n_t = g.nblk[t];
            const int64_t o_row_base = (p.cu_q != nullptr) ? (int64_t)g.q_row0 : (int64_t)w.bidb * p.sq;
            float* lse_row = p.lse + ((int64_t)w.bidb * p.h + w.bidh) * p.sq;
            uint16_t* o_row = o_base + ((o_row_base + row) * p.h + w.bidh) * D;

            if (n_t == 0) {
                // rows exist but see no key: O = 0, LSE = 0
                if (row < g.sq_b) {
#pragma unroll
                    for (int ch = 0; ch < D / 8; ++ch) *(reinterpret_cast<uint4*>(o_row) + ch) = make_uint4(0, 0, 0, 0);
                    lse_row[row] = 0.f;
                }
                continue;
            }
            int col_limit = g.sk_b - 1;
            if (p.is_causal) col_limit = min(col_limit, row + g.causal_off);
            float m_ref = -INFINITY, l_run = 0.f;

            for (int j = 0; j < n_t; ++j) {
                const int n0 = j * kBlockN;
                mbar_wait(&bar_s_full[t], (its + j) & 1);
                tc_fence_after();
                if (r_in_tile == 0) FA_TRACE_EVENT(t, its + j, 0);
                float s[kBlockN];
#pragma unroll
                for (int c = 0; c < kBlockN / 32; ++c)
                    tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
                tmem_wait_ld();
                if (r_in_tile == 0) FA_TRACE_EVENT(t, its + j, 1);
                const bool need_mask = (n0 + kBlockN > g.sk_b) || (p.is_causal && (n0 + kBlockN - 1 > mt + g.causal_off));
                if (need_mask) {
                    const int lim = col_limit - n0;
#pragma unroll
                    for (int c = 0; c < kBlockN; ++c)
                        if (c > lim) s[c] = -INFINITY;
                }
                softmax_step<D, kBf16, kEmu>(s, j == 0, c2, p.inv_scale_log2, m_ref, l_run, tS, tO, &bar_p_full[2 * t],
                                             &bar_p_full[2 * t + 1]);
                if (r_in_tile == 0) FA_TRACE_EVENT(t, its + j, 4);
            }
            its += n_t;

            // ---- epilogue: O_t / l -> 16 bit -> staging tile (128B-swizzled, the TMA layout) -> TMA store; LSE ----
            // (a direct TMEM -> register -> global epilogue was measured first: 16 strided 16-byte stores per thread keep
            //  the LSU busy for 3-5k cycles per tile and delay the mbarrier traffic of the next work item behind them)
            mbar_wait(&bar_o_full[t], nitem & 1);
            tc_fence_after();
            if (r_in_tile == 0) FA_TRACE_EVENT(t, its - 1, 5);
            const bool row_empty = (m_ref == -INFINITY) || !(l_run > 0.f);   // no visible key: O = 0, LSE = 0
            const float inv_l = row_empty ? 0.f : (1.f / l_run);
            if ((warp & 3) == 0) {       // take the staging tile (the other warpgroup may hold it).  The whole warp spins
                int got;                 // together: bar.sync below is warp-aligned, a lone spinning lane would be undefined
                do {
                    got = 0;
                    if (lane == 0) got = (atomicCAS(stage_lock, 0, 1) == 0);
                    got = __shfl_sync(0xffffffffu, got, 0);
                    if (!got) __nanosleep(32);
                } while (!got);
            }
            named_bar_sync(1 + t, kBlockM);
            const uint32_t stage = smem_u32(sStage);
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
                uint32_t o[32];
                tmem_ld32(tO + c * 32, o);
                tmem_wait_ld();
                if (c == D / 32 - 1) {          // O_t is in registers: the next item's first P V may overwrite it
                    tc_fence_before();
                    mbar_arrive(&bar_o_empty[t]);
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 v;
                    v.x = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 0]) * inv_l, __uint_as_float(o[q4 * 8 + 1]) * inv_l);
                    v.y = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 2]) * inv_l, __uint_as_float(o[q4 * 8 + 3]) * inv_l);
                    v.z = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 4]) * inv_l, __uint_as_float(o[q4 * 8 + 5]) * inv_l);
                    v.w = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 6]) * inv_l, __uint_as_float(o[q4 * 8 + 7]) * inv_l);
                    const int chunk = c * 4 + q4;             // 16-byte chunk of the row; 8 chunks per 64-column slab
                    sts128u(stage + (chunk >> 3) * L::kSlab + r_in_tile * 128 + (((chunk & 7) ^ (r_in_tile & 7)) << 4), v);
                }
            }
            if (row < g.sq_b) lse_row[row] = row_empty ? 0.f : (m_ref * p.scale + logf(l_run));
            fence_proxy_async_smem();                          // generic-proxy writes -> visible to the TMA engine
            named_bar_sync(1 + t, kBlockM);
            const bool whole_tile = (mt + kBlockM <= g.sq_b) || (p.cu_q == nullptr);   // dense: TMA clips rows >= seqlen_q itself
            if (whole_tile) {
                if (r_in_tile == 0) {
#pragma unroll
                    for (int sl = 0; sl < kSlabs; ++sl)
                        tma_store_4d(&tmO, sStage + sl * L::kSlab, sl * 64, w.bidh, g.q_row0 + mt, g.tma_b);
                    tma_store_commit();
                    tma_store_wait_read<0>();                  // staging tile has been read; global writes complete later
                    __threadfence_block();
                    atomicExch(stage_lock, 0);
                }
                __syncwarp();
            } else {
                // ragged varlen tail: a TMA box would spill into the next sequence -> predicated coalesced stores
                constexpr int kChunksPerRow = D / 8;
                for (int idx = r_in_tile; idx < kBlockM * kChunksPerRow; idx += kBlockM) {
                    const int rr = idx / kChunksPerRow, ch = idx % kChunksPerRow;
                    if (mt + rr < g.sq_b) {
                        const uint4 v = lds128u(stage + (ch >> 3) * L::kSlab + rr * 128 + (((ch & 7) ^ (rr & 7)) << 4));
                        *(reinterpret_cast<uint4*>(o_base + ((o_row_base + mt + rr) * p.h + w.bidh) * D) + ch) = v;
                    }
                }
                named_bar_sync(1 + t, kBlockM);
                if (r_in_tile == 0) atomicExch(stage_lock, 0);
                __syncwarp();
            }
            if (r_in_tile == 0) FA_TRACE_EVENT(t, its - 1, 6);
            ++nitem;
        }
        if (r_in_tile == 0) tma_store_wait<0>();   // all bulk stores of this thread have landed before the CTA retires
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int D, bool kBf16, int kEmu>
int launch_fwd_persistent(const fa_fwd_params* p, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, FwdParams kp,
                          cudaStream_t stream) {
    using L = FwdSmemP<D>;
    // tensor map of the output (same 4-D view as Q) for the epilogue's TMA store
    CUtensorMap to;
    {
        const bool varlen = p->cu_seqlens_q != nullptr;
        const uint64_t Dd = (uint64_t)p->d, rows_q = varlen ? (uint64_t)p->total_q : (uint64_t)p->seqlen_q, nb = varlen ? 1 : (uint64_t)p->b;
        const uint64_t dims[4] = {Dd, (uint64_t)p->h, rows_q, nb};
        const uint64_t str[3] = {Dd * 2, (uint64_t)p->h * Dd * 2, rows_q * (uint64_t)p->h * Dd * 2};
        const uint32_t box[4] = {64, 1, (uint32_t)kBlockM, 1};
        const int rc = encode_tmap_4d(&to, p->o, p->dtype == FA_DTYPE_BF16, dims, str, box);
        if (rc != FA_OK) return rc;
    }
    auto kern = flash_fwd_kernel_sm100_persistent<D, kBf16, kEmu>;
    static bool attr_set = false;
    static int num_sms = 0;
    if (!attr_set) {
        FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytesP));
        int dev = 0;
        FA_CUDA_CHECK(cudaGetDevice(&dev));
        FA_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_set = true;
    }
    const TileSched ts = make_tile_sched(p);
    const int grid = ts.total < num_sms ? ts.total : num_sms;
    kern<<<grid, 384, L::kBytesP, stream>>>(tq, tk, tv, to, kp, ts);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return FA_OK;
}

// explicit instantiations used by launch_fwd_sm100
#define FA_INST(D, B, E)                                                                                              \
    template int launch_fwd_persistent<D, B, E>(const fa_fwd_params*, const CUtensorMap&, const CUtensorMap&,            \
                                                const CUtensorMap&, FwdParams, cudaStream_t);
FA_INST(128, true, 0) FA_INST(128, true, 1) FA_INST(128, true, 2)
FA_INST(128, false, 0) FA_INST(128, false, 1) FA_INST(128, false, 2)
FA_INST(64, true, 0) FA_INST(64, false, 0)
#undef FA_INST

}  // namespace fa100
