// flash_fwd_p4_sm100.cu — persistent forward with FOUR softmax warpgroups (head_dim 128).
//
// Why: in flash_fwd_persist_sm100.cu one warpgroup (one warp per SM sub-partition) owns a whole 128x128 score tile.
// The exponentials of a tile need 1024 cycles of the SM's MUFU units (16 ex2/clk), exactly as long as the tile's two
// MMAs, and the tile's dependency loop is  S_t -> softmax_t -> P_t V -> S_t' ;  whenever only one of the two tiles is in
// its softmax phase a single warp per sub-partition cannot keep the MUFU pipe full (measured: 1430 cycles per row-tile
// with one warp, 1050-1150 with two; profiles/r01_ubench_softmax_pipes.log), which stretches the loop to ~2900 cycles
// per pair of tiles against 2048 of MMA work.  Here every score row is split between two threads of two different
// warpgroups (columns [0,64) and [64,128)), so each tile's softmax always has two warps per sub-partition and the MUFU
// pipe is saturated even when the tiles' phases do not overlap.
//
// Roles (640 threads): warpgroup 2t+hh = softmax of tile t, column half hh (104 registers/thread);
//                      warp 16 = MMA issuer, warp 17 = TMA producer (warpgroup 4, 64 registers/thread).
// Register budget: five warps per SM sub-partition launch with 96 registers each; setmaxnreg.inc can only draw what
// setmaxnreg.dec released (the launch-time slack of the register file is NOT in the pool: a 112/64 split deadlocks),
// so warpgroup 4 gives up 32 per thread and each of the four softmax warps of a sub-partition takes 8.
// TMEM plan is the one of the two-warpgroup kernel (S0 S1 O0 O1, 128 columns each); P_t half hh overwrites the first
// 32 columns of ITS OWN half of S_t (columns 64hh .. 64hh+32), so a thread can re-read its scores from TMEM in the
// rare rescale path and never has to keep them live in registers.
// The two threads of a row agree on the running reference max through a 4-byte shared-memory slot each and a
// 64-thread named barrier per warp pair; the exponentials still start speculatively with the old reference.
#include <type_traits>

// Two experiments on the item-to-item transition (clock64 trace: ~3600 cycles from a tile's last P to its first S of the
// next item, against 1260 inside an item).  Both were measured neutral (+-1 % at S1k / C2 / C3 / C4,
// profiles/r01s2_run19.log, r01s2_run20.log), so the simpler configuration stays the default:
//   FA_P4_STAGING2  1: three K/V ring stages + one O staging tile PER query tile, TMA stores issued by helper warps 18/19
//                   0: four ring stages + one shared staging tile behind a lock, store issued (and awaited) by a softmax thread
//   FA_P4_QPREFETCH 1: the TMA producer fetches the next item's Q tiles from inside its waits for K/V ring slots
#ifndef FA_P4_STAGING2
#define FA_P4_STAGING2 0
#endif
#ifndef FA_P4_QPREFETCH
#define FA_P4_QPREFETCH 0
#endif
#ifndef FA_P4_SUMVOTE
#define FA_P4_SUMVOTE 1   // 0: per-element row max + vote (A/B builds); 1: vote on the tile sum, max only when it trips
#endif

#include "flash_fwd_common.cuh"

namespace fa100 {

namespace {
constexpr int D = 128;
constexpr int kThreadsP4 = 640;
using L = FwdSmemP<D>;
constexpr int kOffXch = L::kOffBarP;                  // float [2 tiles][2 halves][128 rows], 1024-byte aligned: peer slot = own ^ 512
constexpr int kOffBars = kOffXch + 2 * 2 * kBlockM * 4;
constexpr int kNeedP4 = kOffBars + 256;
constexpr int kBytesP4 = 232448;                      // everything an SM has (227 KB); the slack absorbs a base that is not 1024-byte aligned
static_assert(kNeedP4 + 512 <= kBytesP4, "shared memory budget");
}  // namespace

template <bool kBf16, int kEmu>
__global__ void __launch_bounds__(kThreadsP4, 1)
flash_fwd_kernel_sm100_p4(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                          const FwdParams p, const TileSched ts) {
    constexpr int kSlabs = D / 64;
    constexpr int kStages = FA_P4_STAGING2 ? 3 : L::kKvStages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int wg = warp >> 2;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    if (smem - smem_raw > kBytesP4 - kNeedP4) { asm volatile("trap;"); }   // cannot happen with a >= 256-byte aligned window
    uint8_t* sQ = smem + L::kOffQ;
    uint8_t* sKV = smem + L::kOffKV;
    // with two staging tiles the K/V ring gives up one slot: same total (Q 64 KB + ring + staging = 224 KB)
    uint8_t* sStage = smem + L::kOffKV + kStages * L::kTile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
    uint64_t* bar_q_full = bars;                      // [2]  Q_t landed
    uint64_t* bar_q_empty = bars + 2;                 // [2]  last S_t MMA of the item retired
    uint64_t* bar_kv_full = bars + 4;                 // [kStages]
    uint64_t* bar_kv_empty = bars + 4 + kStages;
    uint64_t* bar_s_full = bars + 4 + 2 * kStages;    // [2]
    uint64_t* bar_p_full = bar_s_full + 2;            // [2 tiles][2 halves], 128 arrivals each (one warpgroup)
    uint64_t* bar_o_full = bar_p_full + 4;            // [2]  last P V of the item retired
    uint64_t* bar_o_empty = bar_o_full + 2;           // [2]  epilogue has O_t in registers (256 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o_empty + 2);
    int* stage_lock = reinterpret_cast<int*>(tmem_slot + 1);   // FA_P4_STAGING2 == 0: the two tiles' epilogues share one staging tile
    uint64_t* bar_stage_free = bars + 28;             // [2] FA_P4_STAGING2: the TMA store of tile t has read its staging tile
    const uint32_t xch = smem_u32(smem + kOffXch);

    if (warp == 16) {
        if (lane == 0) {
            *stage_lock = 0;
            for (int t = 0; t < 2; ++t) {
                mbar_init(&bar_q_full[t], 1); mbar_init(&bar_q_empty[t], 1);
                mbar_init(&bar_s_full[t], 1);
                mbar_init(&bar_p_full[2 * t], kBlockM); mbar_init(&bar_p_full[2 * t + 1], kBlockM);
                mbar_init(&bar_o_full[t], 1); mbar_init(&bar_o_empty[t], 2 * kBlockM);
            }
            for (int i = 0; i < kStages; ++i) { mbar_init(&bar_kv_full[i], 1); mbar_init(&bar_kv_empty[i], 1); }
            mbar_init(&bar_stage_free[0], 1); mbar_init(&bar_stage_free[1], 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
        tmem_relinquish();
    } else if (warp == 17 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wg == 4) {
        setmaxnreg_dec<64>();
        if (warp == 17) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                int kv_i = 0;            // running K/V ring index
                int nq[2] = {0, 0};      // Q_t loads so far
                bool q_ahead[2] = {false, false};   // Q_t of the item being started was already issued during the previous item
                for (int n = blockIdx.x; n < ts.total; n += gridDim.x) {
                    const WorkItem w = decode_item(ts, n, p.h, p.is_causal != 0);
                    const ItemGeom g = item_geom(p, w);
                    if (g.skip || g.n_blocks == 0) continue;
                    const int bidh_k = w.bidh / p.hratio;
                    // The next item's Q tiles do not go through the K/V ring: as soon as the last S_t MMA of this item has
                    // retired (bar_q_empty[t]) they are fetched, from inside the waits for ring slots below.  Without this the
                    // Q loads were only issued once every K/V load of the current item was out, and the first S of the next
                    // item arrived ~1500 cycles after the softmax warps were ready for it (clock64 trace).
                    const int nn = n + gridDim.x;
                    bool nvalid = FA_P4_QPREFETCH && nn < ts.total;
                    WorkItem w2 = w;
                    ItemGeom g2 = g;
                    if (nvalid) {
                        w2 = decode_item(ts, nn, p.h, p.is_causal != 0);
                        g2 = item_geom(p, w2);
                        nvalid = !(g2.skip || g2.n_blocks == 0);
                    }
                    bool q_next[2] = {false, false};
                    bool own_q_done = false;
                    auto issue_q = [&](const WorkItem& wi, const ItemGeom& gi, int t) {
                        mbar_arrive_expect_tx(&bar_q_full[t], L::kTile);
                        for (int s = 0; s < kSlabs; ++s)
                            tma_load_4d(sQ + t * L::kTile + s * L::kSlab, &tmQ, &bar_q_full[t], s * 64, wi.bidh,
                                        gi.q_row0 + gi.m0 + t * kBlockM, gi.tma_b);
                        ++nq[t];
                    };
                    auto try_prefetch_q = [&]() {
                        if (!nvalid || !own_q_done) return;
#pragma unroll
                        for (int t = 0; t < 2; ++t)
                            if (!q_next[t] && g2.nblk[t] > 0 && mbar_try_wait(&bar_q_empty[t], (nq[t] & 1) ^ 1)) {
                                issue_q(w2, g2, t);
                                q_next[t] = true;
                            }
                    };
                    auto load_kv = [&](const CUtensorMap* tm, int j) {
                        const int slot = kv_i % kStages;
                        while (!mbar_try_wait(&bar_kv_empty[slot], ((kv_i / kStages) & 1) ^ 1)) try_prefetch_q();
                        mbar_arrive_expect_tx(&bar_kv_full[slot], L::kTile);
                        for (int s = 0; s < kSlabs; ++s)
                            tma_load_4d(sKV + slot * L::kTile + s * L::kSlab, tm, &bar_kv_full[slot], s * 64, bidh_k,
                                        g.k_row0 + j * kBlockN, g.tma_b);
                        ++kv_i;
                    };
                    auto load_q = [&](int t) {
                        if (g.nblk[t] == 0) return;
                        if (q_ahead[t]) return;                  // issued while the previous item was still running
                        mbar_wait(&bar_q_empty[t], (nq[t] & 1) ^ 1);
                        issue_q(w, g, t);
                    };
                    load_q(0);
                    load_kv(&tmK, 0);
                    load_q(1);
                    own_q_done = true;
                    load_kv(&tmV, 0);
                    for (int j = 1; j < g.n_blocks; ++j) {
                        load_kv(&tmK, j);
                        load_kv(&tmV, j);
                    }
                    try_prefetch_q();
                    q_ahead[0] = q_next[0]; q_ahead[1] = q_next[1];
                }
            }
        } else if (warp == 16) {
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
                auto issue_pv = [&](int t, int j, int half) {  // O_t += P_t[:, half] V_j[half];  P half at S_t + 64 half
                    if (leader) {
                        const uint32_t va = v_lo + kv_slot(2 * j + 1) * kTile16;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_ts(tm + kTmemO0 + t * 128, tm + kTmemS0 + t * 128 + half * 64 + kk * 8,
                                    desc_make(va + (half * 4 + kk) * (2048 >> 4), kDescHiK), idesc_pv, (j > 0 || half > 0 || kk > 0));
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
#if FA_P4_STAGING2
        else {
            // ===================== warps 18 / 19: TMA store of tile slot 0 / 1 =====================
            // The softmax warpgroups only write the staging tile and arrive on a named barrier; waiting for the bulk
            // engine to read 32 KB of shared memory (~1800 cycles, measured) is this warp's job, not theirs.
            const int t = warp - 18;
            for (int n = blockIdx.x; n < ts.total; n += gridDim.x) {
                const WorkItem w = decode_item(ts, n, p.h, p.is_causal != 0);
                const ItemGeom g = item_geom(p, w);
                if (g.skip) continue;
                const int mt = g.m0 + t * kBlockM;
                if (mt >= g.sq_b || g.nblk[t] == 0) continue;
                const bool whole_tile = (mt + kBlockM <= g.sq_b) || (p.cu_q == nullptr);
                if (!whole_tile) continue;                         // ragged varlen tail: stored by the softmax threads themselves
                named_bar_sync(11 + t, 2 * kBlockM + 32);          // staging tile t written and fenced by its 256 threads
                if (lane == 0) {
#pragma unroll
                    for (int sl = 0; sl < kSlabs; ++sl)
                        tma_store_4d(&tmO, sStage + t * L::kTile + sl * L::kSlab, sl * 64, w.bidh, g.q_row0 + mt, g.tma_b);
                    tma_store_commit();
                    tma_store_wait_read<0>();
                    mbar_arrive(&bar_stage_free[t]);
                }
                __syncwarp();
            }
            if (lane == 0) tma_store_wait<0>();                    // all bulk stores have landed before the CTA retires
        }
#endif
    } else {
        // ========== softmax warpgroups: warpgroup 2t+hh owns columns [64hh, 64hh+64) of tile slot t, one thread per row ==========
        setmaxnreg_inc<104>();
        const int t = wg >> 1;
        const int hh = wg & 1;
        const int wq = warp & 3;                         // TMEM lane quadrant = SM sub-partition
        const int r_in_tile = tid & 127;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        const uint32_t tS = tmem_base + lane_base + kTmemS0 + t * 128 + hh * 64;   // own scores; own P half goes to the same place
        const uint32_t tO = tmem_base + lane_base + kTmemO0 + t * 128 + hh * 64;   // own half of the O row
        const uint32_t x_own = xch + ((t * 2 + hh) * kBlockM + r_in_tile) * 4;
        const uint32_t pair_bar = 1 + t * 4 + wq;       // named barriers 1..8: the two warps that share 32 rows (64 threads)
        const uint32_t tile_bar = 9 + t;                // named barriers 9, 10: the two warpgroups of a tile (256 threads)
        uint16_t* o_base = reinterpret_cast<uint16_t*>(p.o);
        const float c2 = p.scale_log2;
        const float inv_c2 = p.inv_scale_log2;
        int its = 0;       // S_t steps so far
        int nitem = 0;     // items with keys finished by this slot
#if FA_P4_STAGING2
        int nstore = 0;    // TMA stores of this slot handed to the helper warp so far
#endif

        // The 64 scores of a thread are walked in four chunks of 16 columns (tcgen05.ld x16), the next chunk in flight
        // while the current one is processed: only 32 score registers are ever live next to the 32 packed P words, which
        // is what lets this role fit 104 registers without local-memory traffic in the loop (L1 is ~0 KB here: shared
        // memory takes all of it, a spill costs an L2 round trip).
        // P = 2^(s*c2 + neg); kEmu of every 4 column pairs go through the Cody-Waite + degree-3 polynomial path (FMA pipe).
        auto exp_chunk = [&](const float (&s)[16], const float neg, uint32_t* pk8, float2& sum) {
            const float2 c2v = make_float2(c2, c2);
            const float2 negv = make_float2(neg, neg);
            // kEmu -> emulated pairs of every 8 (spread evenly): 4 -> {0}, 1 -> {0,4}, 3 -> {0,3,6}, 2 -> {0,2,4,6}
            constexpr int kEmu8 = kEmu == 1 ? 2 : kEmu == 2 ? 4 : kEmu == 3 ? 3 : kEmu == 4 ? 1 : 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 pp;
                if (((i * kEmu8) & 7) < kEmu8) {
                    const float2 magic = make_float2(12582912.f, 12582912.f);       // 1.5 * 2^23
                    // Clamp the argument to [-125, 126]: below, the result would be subnormal; above, the exponent add
                    // wraps around (2^141 came out as -2^-116, the tile-sum vote did not trip and a whole key tile was lost:
                    // tests/test_parity_gpu.py::test_rescale_path_scores_growing_along_the_keys).  At 126 the emulated value
                    // is ~2^126: the sum trips the vote exactly like the +inf of MUFU.EX2 does.
                    const float s_floor = (-125.f - neg) * inv_c2, s_ceil = (126.f - neg) * inv_c2;
                    const float2 x = __ffma2_rn(make_float2(fminf(fmaxf(s[2 * i], s_floor), s_ceil),
                                                            fminf(fmaxf(s[2 * i + 1], s_floor), s_ceil)), c2v, negv);
                    const float2 tt = __fadd2_rn(x, magic);                                   // low mantissa bits = rint(x)
                    const float2 nnf = __ffma2_rn(tt, make_float2(-1.f, -1.f), magic);        // -rint(x), exact
                    const float2 f = __fadd2_rn(x, nnf);                                      // x - rint(x) in [-0.5, 0.5]
                    float2 pl = __ffma2_rn(make_float2(0.05517115816473961f, 0.05517115816473961f), f,
                                           make_float2(0.2426101416349411f, 0.2426101416349411f));
                    pl = __ffma2_rn(pl, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
                    pl = __ffma2_rn(pl, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
                    pp = make_float2(__uint_as_float(__float_as_uint(pl.x) + (__float_as_uint(tt.x) << 23)),
                                     __uint_as_float(__float_as_uint(pl.y) + (__float_as_uint(tt.y) << 23)));
                } else {
                    const float2 x = __ffma2_rn(make_float2(s[2 * i], s[2 * i + 1]), c2v, negv);
                    pp = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                }
                sum = __fadd2_rn(sum, pp);
                pk8[i] = pack2<kBf16>(pp.x, pp.y);
            }
        };
        auto mask_chunk = [&](float (&s)[16], const int lim_c) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c > lim_c) s[c] = -INFINITY;
        };
        auto max_chunk = [&](const float (&s)[16], float mx) -> float {
            float ma = fmaxf(s[0], s[1]), mb = fmaxf(s[2], s[3]);
#pragma unroll
            for (int c = 4; c < 16; c += 4) {
                ma = fmaxf(ma, fmaxf(s[c], s[c + 1]));
                mb = fmaxf(mb, fmaxf(s[c + 2], s[c + 3]));
            }
            return fmaxf(mx, fmaxf(ma, mb));
        };
        // one pass over the thread's 64 scores: kExp -> exponentials into pk / sum; kMax -> running max into mx
        auto walk_m = [&](auto do_exp, auto do_max, auto do_mask, const int lim, const float neg, uint32_t (&pk)[32],
                          float2& sum, float& mx) {
            float sa[16], sb[16];
            tmem_ld16(tS, *reinterpret_cast<uint32_t(*)[16]>(&sa[0]));
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float (&cur)[16] = (c & 1) ? sb : sa;
                float (&nxt)[16] = (c & 1) ? sa : sb;
                if (c < 3) tmem_ld16(tS + 16 * (c + 1), *reinterpret_cast<uint32_t(*)[16]>(&nxt[0]));
                if constexpr (decltype(do_mask)::value) mask_chunk(cur, lim - 16 * c);
                if constexpr (decltype(do_max)::value) mx = max_chunk(cur, mx);
                if constexpr (decltype(do_exp)::value) exp_chunk(cur, neg, &pk[8 * c], sum);
                if (c < 3) tmem_wait_ld();
            }
        };
        using yes = std::true_type;
        using no = std::false_type;
        // masked and unmasked tiles get separate straight-line copies (no per-chunk branches: the scheduler can overlap
        // the tail of one chunk with the head of the next)
        auto walk = [&](auto do_exp, auto do_max, const bool need_mask, const int lim, const float neg, uint32_t (&pk)[32],
                        float2& sum, float& mx) {
            if (need_mask) walk_m(do_exp, do_max, yes{}, lim, neg, pk, sum, mx);
            else walk_m(do_exp, do_max, no{}, lim, neg, pk, sum, mx);
        };

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

            if (n_t == 0) {
                // rows exist but see no key: O = 0, LSE = 0 (each half clears its 64 columns)
                if (row < g.sq_b) {
                    uint16_t* o_row = o_base + ((o_row_base + row) * p.h + w.bidh) * D + hh * 64;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) *(reinterpret_cast<uint4*>(o_row) + ch) = make_uint4(0, 0, 0, 0);
                    if (hh == 0) lse_row[row] = 0.f;
                }
                continue;
            }
            int col_limit = g.sk_b - 1;
            if (p.is_causal) col_limit = min(col_limit, row + g.causal_off);
            float m_ref = -INFINITY, l_run = 0.f;

            for (int j = 0; j < n_t; ++j) {
                const int n0 = j * kBlockN;
                const bool need_mask = (n0 + kBlockN > g.sk_b) || (p.is_causal && (n0 + kBlockN - 1 > mt + g.causal_off));
                const int lim = col_limit - n0 - hh * 64;
                mbar_wait(&bar_s_full[t], (its + j) & 1);
                tc_fence_after();
                if (r_in_tile == 0 && hh == 0) FA_TRACE_EVENT(t, its + j, 0);
                uint32_t pk[32];
                float2 sum = make_float2(0.f, 0.f);
                float mx = -INFINITY;
                if (j == 0) {       // no reference yet: exact row max (both halves) before the exponentials
                    walk(no{}, yes{}, need_mask, lim, 0.f, pk, sum, mx);
                    sts32f(x_own, mx);
                    named_bar_sync(pair_bar, 64);
                    float mp;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mp) : "r"(x_own ^ (kBlockM * 4)));
                    m_ref = fmaxf(mx, mp);
                    const float neg = (m_ref == -INFINITY) ? 0.f : -m_ref * c2;
                    walk(yes{}, no{}, need_mask, lim, neg, pk, sum, mx);
                    if (r_in_tile == 0 && hh == 0) FA_TRACE_EVENT(t, its + j, 2);
                } else {
                    // speculative: exponentials with the old reference; the vote below confirms the reference (or sends
                    // the warp pair through the rescale path)
                    float neg = (m_ref == -INFINITY) ? 0.f : -m_ref * c2;
#if FA_P4_SUMVOTE
                    // No per-element max: every P is non-negative, so "tile sum <= 2^9" proves that no exponent exceeded 9
                    // (the lazy-rescale bound); only a row whose sum is larger (or inf / NaN) pays for a true max.
                    walk(yes{}, no{}, need_mask, lim, neg, pk, sum, mx);
                    if (r_in_tile == 0 && hh == 0) FA_TRACE_EVENT(t, its + j, 2);
                    const float hs = sum.x + sum.y;
                    sts32f(x_own, hs);
                    named_bar_sync(pair_bar, 64);
                    float hp;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hp) : "r"(x_own ^ (kBlockM * 4)));
                    bool need = !(hs + hp <= 512.f) || (m_ref == -INFINITY);   // (a row without a reference yet must look at its max)
                    if (__any_sync(0xffffffffu, need)) {
                        float mxo = -INFINITY;
                        walk(no{}, yes{}, need_mask, lim, neg, pk, sum, mxo);
                        named_bar_sync(pair_bar, 64);          // both threads of the row have read the sums: slots reusable
                        sts32f(x_own, mxo);
                        named_bar_sync(pair_bar, 64);
                        float mp2;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mp2) : "r"(x_own ^ (kBlockM * 4)));
                        mx = fmaxf(mxo, mp2);
                        need = need && (mx > m_ref);
                    }
#else
                    walk(yes{}, yes{}, need_mask, lim, neg, pk, sum, mx);
                    if (r_in_tile == 0 && hh == 0) FA_TRACE_EVENT(t, its + j, 2);
                    sts32f(x_own, mx);
                    named_bar_sync(pair_bar, 64);
                    float mp;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mp) : "r"(x_own ^ (kBlockM * 4)));
                    mx = fmaxf(mx, mp);
                    const bool need = (mx - m_ref) * c2 > kRescaleThreshold;   // reference moves by more than 2^8
#endif
                    if (__any_sync(0xffffffffu, need)) {
                        // slow path (both warps of the pair take it together: they see the same 32 row maxima):
                        // rescale the own half of O_t and the running sum, redo the exponentials with the new reference
                        float alpha = 1.f;
                        if (need) {
                            alpha = fast_exp2((m_ref - mx) * c2);
                            m_ref = mx;
                            l_run *= alpha;
                        }
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint32_t o[32];
                            tmem_ld32(tO + c * 32, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st32(tO + c * 32, o);
                        }
                        tmem_wait_st();
                        neg = (m_ref == -INFINITY) ? 0.f : -m_ref * c2;
                        sum = make_float2(0.f, 0.f);
                        walk(yes{}, no{}, need_mask, lim, neg, pk, sum, mx);   // own S columns are intact: own P not stored yet
                        // P_t V of this step accumulates into ALL of O_t: neither half may release its P before both
                        // halves of the row block have finished rescaling
                        tc_fence_before();
                        named_bar_sync(pair_bar, 64);
                        tc_fence_after();
                    }
                }
                if (r_in_tile == 0 && hh == 0) FA_TRACE_EVENT(t, its + j, 3);
                tmem_st32(tS, pk);
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(&bar_p_full[2 * t + hh]);
                l_run += sum.x + sum.y;
                if (r_in_tile == 0 && hh == 0) FA_TRACE_EVENT(t, its + j, 4);
            }
            its += n_t;

            // ---- epilogue: O_t / l -> 16 bit -> staging tile (128B-swizzled, the TMA layout; half hh = slab hh) -> TMA store ----
            mbar_wait(&bar_o_full[t], nitem & 1);
            tc_fence_after();
            if (r_in_tile == 0 && hh == 0) FA_TRACE_EVENT(t, its - 1, 5);
            sts32f(x_own, l_run);
            named_bar_sync(pair_bar, 64);
            float l_peer;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l_peer) : "r"(x_own ^ (kBlockM * 4)));
            const float l_tot = l_run + l_peer;
            const bool row_empty = (m_ref == -INFINITY) || !(l_tot > 0.f);   // no visible key: O = 0, LSE = 0
            const float inv_l = row_empty ? 0.f : (1.f / l_tot);
#if FA_P4_STAGING2
            if (nstore > 0) mbar_wait(&bar_stage_free[t], (nstore - 1) & 1);   // the previous store of this slot has read the tile
            uint8_t* sStageT = sStage + t * L::kTile;
#else
            if (hh == 0 && wq == 0) {    // take the staging tile (the other tile's epilogue may hold it); the whole warp spins
                int got;                 // together: bar.sync below is warp-aligned
                do {
                    got = 0;
                    if (lane == 0) got = (atomicCAS(stage_lock, 0, 1) == 0);
                    got = __shfl_sync(0xffffffffu, got, 0);
                    if (!got) __nanosleep(32);
                } while (!got);
            }
            named_bar_sync(tile_bar, 2 * kBlockM);
            uint8_t* sStageT = sStage;
#endif
            if (r_in_tile == 0 && hh == 0) FA_TRACE_EVENT(t, its - 1, 1);
            const uint32_t stage = smem_u32(sStageT) + hh * L::kSlab;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t o[32];
                tmem_ld32(tO + c * 32, o);
                tmem_wait_ld();
                if (c == 1) {                   // O_t is in registers: the next item's first P V may overwrite it
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
                    const int chunk = c * 4 + q4;             // 16-byte chunk of this half's 128-byte slab row
                    sts128u(stage + r_in_tile * 128 + ((chunk ^ (r_in_tile & 7)) << 4), v);
                }
            }
            if (hh == 0 && row < g.sq_b) lse_row[row] = row_empty ? 0.f : (m_ref * p.scale + logf(l_tot));
            fence_proxy_async_smem();                          // generic-proxy writes -> visible to the TMA engine
            const bool whole_tile = (mt + kBlockM <= g.sq_b) || (p.cu_q == nullptr);   // dense: TMA clips rows >= seqlen_q itself
#if FA_P4_STAGING2
            if (whole_tile) {
                named_bar_arrive(11 + t, 2 * kBlockM + 32);    // hand the tile to helper warp 18 + t and move on
                ++nstore;
            } else {
                named_bar_sync(tile_bar, 2 * kBlockM);
            }
#else
            named_bar_sync(tile_bar, 2 * kBlockM);
#endif
            if (r_in_tile == 0 && hh == 0) FA_TRACE_EVENT(t, its - 1, 7);
            if (whole_tile) {
                if (!FA_P4_STAGING2 && hh == 0 && r_in_tile == 0) {
#pragma unroll
                    for (int sl = 0; sl < kSlabs; ++sl)
                        tma_store_4d(&tmO, sStageT + sl * L::kSlab, sl * 64, w.bidh, g.q_row0 + mt, g.tma_b);
                    tma_store_commit();
                    tma_store_wait_read<0>();                  // staging tile has been read; global writes complete later
                    __threadfence_block();
                    atomicExch(stage_lock, 0);
                }
                __syncwarp();
            } else {
                // ragged varlen tail: a TMA box would spill into the next sequence -> predicated coalesced stores
                constexpr int kChunksPerRow = D / 8;
                const uint32_t stage0 = smem_u32(sStageT);
                for (int idx = hh * kBlockM + r_in_tile; idx < kBlockM * kChunksPerRow; idx += 2 * kBlockM) {
                    const int rr = idx / kChunksPerRow, ch = idx % kChunksPerRow;
                    if (mt + rr < g.sq_b) {
                        const uint4 v = lds128u(stage0 + (ch >> 3) * L::kSlab + rr * 128 + (((ch & 7) ^ (rr & 7)) << 4));
                        *(reinterpret_cast<uint4*>(o_base + ((o_row_base + mt + rr) * p.h + w.bidh) * D) + ch) = v;
                    }
                }
                named_bar_sync(tile_bar, 2 * kBlockM);
                if (!FA_P4_STAGING2 && hh == 0 && r_in_tile == 0) atomicExch(stage_lock, 0);
                __syncwarp();
            }
            if (r_in_tile == 0 && hh == 0) FA_TRACE_EVENT(t, its - 1, 6);
            ++nitem;
        }
        if (!FA_P4_STAGING2 && hh == 0 && r_in_tile == 0) tma_store_wait<0>();   // all bulk stores of this thread have landed before the CTA retires
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <bool kBf16, int kEmu>
int launch_fwd_p4(const fa_fwd_params* p, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, FwdParams kp,
                  cudaStream_t stream) {
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
    auto kern = flash_fwd_kernel_sm100_p4<kBf16, kEmu>;
    static bool attr_set = false;
    static int num_sms = 0;
    if (!attr_set) {
        FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytesP4));
        int dev = 0;
        FA_CUDA_CHECK(cudaGetDevice(&dev));
        FA_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_set = true;
    }
    const TileSched ts = make_tile_sched(p);
    const int grid = ts.total < num_sms ? ts.total : num_sms;
    kern<<<grid, kThreadsP4, kBytesP4, stream>>>(tq, tk, tv, to, kp, ts);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return FA_OK;
}

#define FA_INST(B, E)                                                                                                  \
    template int launch_fwd_p4<B, E>(const fa_fwd_params*, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,  \
                                     FwdParams, cudaStream_t);
FA_INST(true, 0) FA_INST(true, 1) FA_INST(true, 2) FA_INST(true, 3) FA_INST(true, 4)
FA_INST(false, 0) FA_INST(false, 1) FA_INST(false, 2) FA_INST(false, 3) FA_INST(false, 4)
#undef FA_INST

}  // namespace fa100
