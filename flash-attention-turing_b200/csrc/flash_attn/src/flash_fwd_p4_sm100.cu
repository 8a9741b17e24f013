// flash_fwd_p4_sm100.cu — the forward kernel: persistent, warp-specialised, FOUR softmax warpgroups (head_dim 64 / 128).
//
// Replaces the reference's compute_attn_1rowblock (/root/reference/csrc/flash_attn/src/flash_fwd_kernel.h:23-789).
//
// One CTA per SM walks a static list of work items; one item = two 128-row query tiles of one (batch, head), key tiles
// of 128.  TMEM: S0 S1 (fp32 scores, 128 columns each), O0 O1 (fp32 accumulators, head_dim columns each).  Tensor-pipe
// order   S0 S1 | PV0 S0' | PV1 S1' | ...   — in order per issuing thread, so "S_t' complete" implies "PV_t complete".
//
// Roles (640 threads): warpgroup 2t+hh = softmax of tile t, key columns [64hh, 64hh+64), one thread per query row
//                      (104 registers/thread); warp 16 / 17 = MMA issuer of tile 0 / 1 (they pass a token so that the
//                      tiles' blocks of MMAs alternate), warp 18 = TMA producer, warp 19 = TMA store of the O tiles
//                      (64 registers/thread).  A 21st warp does not fit the register file of a sub-partition.
// Why two threads per row: the exponentials of a 128x128 tile need 1024 cycles of the SM's MUFU units (16 ex2/clk), as
// long as the tile's two MMAs, and a tile's dependency loop is S_t -> softmax_t -> P_t V -> S_t'; the two tiles' softmax
// phases mostly alternate, and ONE warp per SM sub-partition cannot keep the MUFU pipe full on its own (measured 1430
// cycles per row-tile against 1050-1150 with two, profiles/r01_ubench_softmax_pipes.log).
//
// Key loop of a softmax thread (details at key_step / spec_step below):
//   * exact step (first step of an item, every step of pass 1): row max of the thread's 64 scores, exchanged with
//     the peer thread of the row through a 4-byte shared-memory slot and a 64-thread named barrier; lazy reference (moves
//     only when the max grows by > 2^8); then the exponentials.
//   * speculative step (all others): exponentials against the reference agreed so far; the two threads of a row publish
//     their step sums and read each other's one step late ("late-agreed reference"); an overflow inside one key tile raises
//     the CTA's retry flag and the CTA walks its items once more with exact steps (pass 1).
//   * P leaves in four 16-column quarters per thread (TMEM store, then one elected arrive per warp half a chunk later);
//     the MMA warp issues P V in the same order, k-steps (q, 4+q) for quarter q.
// P quarter q of half hh overwrites columns [64hh + 8q, +8) of S_t, which have been consumed by then.
#include <atomic>
#include <type_traits>

#include "flash_fwd_common.cuh"

#ifndef FA_P4_PASS1_WIDE
#define FA_P4_PASS1_WIDE 0   // 1: pass 1 loads all four score chunks at once (64 registers in flight) instead of 32 + 16 + 16
#endif

namespace fa100 {

namespace {
constexpr int kThreadsP4 = 640;                        // a 21st warp does not fit: a sub-partition holds 4 x 104 x 32 + 64 x 32 of its 16384 registers
constexpr int kMma1Warp = 17;                          // MMA issuer of tile 1 (warp 16: tile 0)
constexpr int kProducerWarp = 18;                      // TMA producer
                                                       // warp 19: TMA store of the O tiles (one staging tile, shared by both query tiles)
template <int D> struct P4Smem {
    static constexpr int kSlab = kBlockM * 128;          // 64-column slab of a 128-row tile: 16 KB
    static constexpr int kSlabs = D / 64;
    static constexpr int kTile = kSlabs * kSlab;         // one 128 x D tile (32 KB / 16 KB)
    // K/V ring of four tiles (D = 128): V_j, K_j+1 in use, V_j+1 and K_j+2 arriving.  Three slots are not enough: K_j+2 could
    // then only be requested when V_j is released (end of step j) and is needed half a period later — clock64 showed the
    // MMA warp waiting 500-600 cycles per step for K and V, C2 dropped from ~1300 to ~1200 TFLOP/s (profiles/r02_run8.log).
    // So there is ONE O staging tile for both query tiles, guarded by a lock.
    static constexpr int kStages = (D == 128) ? 4 : 8;
    static constexpr int kOffQ = 0;                      // 2 tiles
    static constexpr int kOffKV = 2 * kTile;
    static constexpr int kOffStage = kOffKV + kStages * kTile;   // 16-bit O tile: source of the TMA store
    static constexpr int kOffXch = kOffStage + kTile;    // float [2 tiles][2 halves][128 rows]: peer slot = own ^ 512
    static constexpr int kOffBars = kOffXch + 2 * 2 * kBlockM * 4;
    static constexpr int kNeed = kOffBars + 512;
    static constexpr int kBytes = (D == 128) ? 232448 : kNeed + 1024;   // slack absorbs a base that is not 1024-byte aligned
    static_assert(kNeed + 512 <= kBytes && kBytes <= 232448, "shared memory budget");
};
}  // namespace

template <int D, bool kBf16, int kEmu>
__global__ void __launch_bounds__(kThreadsP4, 1)
flash_fwd_kernel_sm100_p4(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                          const FwdParams p, const TileSched ts) {
    using L = P4Smem<D>;
    constexpr int kSlabs = L::kSlabs;
    constexpr int kStages = L::kStages;
    constexpr int kHalfD = D / 2;                        // O columns owned by one thread (rescale + epilogue)

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int wg = warp >> 2;

    // Only 32-bit addresses in the shared window are kept (never the generic base pointer): see smem_u32 in sm100_ptx.cuh.
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (smem - smem_u32(smem_raw) > (uint32_t)(L::kBytes - L::kNeed)) { asm volatile("trap;"); }   // cannot happen with a >= 512-byte aligned window
    const uint32_t sQ = smem + L::kOffQ;
    const uint32_t sKV = smem + L::kOffKV;
    const uint32_t sStage = smem + L::kOffStage;
    const uint32_t bars = smem + L::kOffBars;             // mbarriers, 8 bytes each: bar_x + 8 i
    const uint32_t bar_q_full = bars;                     // [2]  Q_t landed
    const uint32_t bar_q_empty = bars + 8 * 2;            // [2]  last S_t MMA of the item retired
    const uint32_t bar_kv_full = bars + 8 * 4;            // [kStages]
    const uint32_t bar_kv_empty = bar_kv_full + 8 * kStages;   // [kStages]
    const uint32_t bar_s_full = bar_kv_empty + 8 * kStages;    // [2]
    const uint32_t bar_p = bar_s_full + 8 * 2;            // [2 tiles][4 quarters], 8 arrivals each (one per softmax warp of the tile)
    const uint32_t bar_o_full = bar_p + 8 * 8;            // [2]  last P V of the item retired
    const uint32_t bar_o_empty = bar_o_full + 8 * 2;      // [2]  epilogue has O_t in registers (8 arrivals)
    const uint32_t bar_tok = bar_o_empty + 8 * 2;         // [2]  "tile t's MMA warp may issue its next block"
    const uint32_t tmem_slot = bar_tok + 8 * 2;           // 32-bit words from here on
    const uint32_t stage_lock = tmem_slot + 4;            // the two tiles' epilogues share one staging tile
    const uint32_t retry_flag = stage_lock + 4;           // pass 0: some row's speculative step overflowed (see "late-agreed reference")
    const uint32_t retry_final = stage_lock + 8;          // the flag as published ONCE at the pass boundary: what pass 1 is decided by
    const uint32_t store_desc = stage_lock + 12;          // what the staging tile holds {head, first row, batch}
    const uint32_t xch = smem + L::kOffXch;

    if (warp == 16) {
        if (lane == 0) {
            sts32(stage_lock, 0);
            sts32(retry_flag, 0);
            sts32(retry_final, 0);
            for (int t = 0; t < 2; ++t) {
                mbar_init(bar_q_full + 8 * (t), 1); mbar_init(bar_q_empty + 8 * (t), 1);
                mbar_init(bar_s_full + 8 * (t), 1);
                for (int q = 0; q < 4; ++q) mbar_init(bar_p + 8 * (4 * t + q), 8);
                mbar_init(bar_o_full + 8 * (t), 1); mbar_init(bar_o_empty + 8 * (t), 8);
            }
            // K/V slots are released by the MMA warp(s): with one issuing warp per tile both have to let go
            for (int i = 0; i < kStages; ++i) { mbar_init(bar_kv_full + 8 * (i), 1); mbar_init(bar_kv_empty + 8 * (i), 2); }
            mbar_init(bar_tok + 8 * (0), 1); mbar_init(bar_tok + 8 * (1), 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
        tmem_relinquish();
    } else if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = (uint32_t)lds32(tmem_slot);

    // ---- work list: pass 0 = this CTA's share of the static schedule; pass 1 = the same items once more, with the exact
    // per-step row max, iff some row's speculative exponentials overflowed in pass 0 (one flag per CTA).
    // Round 2 first kept a per-item retry list that the store warp appended to while it checked per-tile flags; a flag
    // consumed late changed the list's length WHILE pass 1 was running, the roles read different item counts and the CTA
    // hung (head_dim 64, fp16, scores of ~N(0, 16^2), three or more items per CTA; found by the random-shape stress,
    // located with cuda-gdb: profiles/r02_run31_gdb_hang.log).  Nothing pass 1 reads may change while it runs: the decision
    // is published once, at the pass boundary, and pass 0's flag is never looked at again. ----
    const int slots_mine = (ts.total > (int)blockIdx.x) ? (ts.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_static = ts.paired ? 2 * slots_mine : slots_mine;      // items of this CTA in the static schedule
    auto static_item = [&](int idx) -> int {                           // idx-th item: slot blockIdx + k gridDim (both halves of a pair)
        return ts.paired ? 2 * ((int)blockIdx.x + (idx >> 1) * (int)gridDim.x) + (idx & 1) : (int)blockIdx.x + idx * (int)gridDim.x;
    };
    auto pass_count = [&](int pass) -> int {
        if (pass == 0) return n_static;
        return lds32(retry_final) ? n_static : 0;
    };
    auto pass_item = [&](int pass, int idx) -> int {
        (void)pass;
        return static_item(idx);
    };
    // every thread of the CTA calls this once, between the passes (each role from its own branch: barrier 0 counts threads)
    auto pass_boundary = [&]() {
        __syncthreads();
        if (tid == 0) sts32(retry_final, lds32(retry_flag));
        __syncthreads();
    };
    const bool exact0 = p.exact != 0;                 // FA_B200_FWD_EXACT=1: no speculation at all (A/B runs, debugging)

    if (wg == 4) {
        setmaxnreg_dec<64>();
        if (warp == kProducerWarp) {
            // ===================== TMA producer =====================
            int kv_i = 0;            // running K/V ring index
            int nq[2] = {0, 0};      // Q_t loads so far
            for (int pass = 0; pass < 2; ++pass) {
              const int cnt = pass_count(pass);
              if (lane == 0) {
                for (int idx = 0; idx < cnt; ++idx) {
                    const int n = pass_item(pass, idx);
                    const WorkItem w = decode_item(ts, n, p.h, p.is_causal != 0);
                    const ItemGeom g = item_geom(p, w);
                    if (g.skip || g.n_blocks == 0) continue;
                    const int bidh_k = w.bidh / p.hratio;
                    auto load_kv = [&](const CUtensorMap* tm, int j) {
                        const int slot = kv_i % kStages;
                        mbar_wait(bar_kv_empty + 8 * (slot), ((kv_i / kStages) & 1) ^ 1);
                        mbar_arrive_expect_tx(bar_kv_full + 8 * (slot), L::kTile);
                        for (int s = 0; s < kSlabs; ++s)
                            tma_load_4d(sKV + slot * L::kTile + s * L::kSlab, tm, bar_kv_full + 8 * (slot), s * 64, bidh_k,
                                        g.k_row0 + j * kBlockN, g.tma_b);
                        ++kv_i;
                    };
                    auto load_q = [&](int t) {
                        if (g.nblk[t] == 0) return;
                        mbar_wait(bar_q_empty + 8 * (t), (nq[t] & 1) ^ 1);
                        mbar_arrive_expect_tx(bar_q_full + 8 * (t), L::kTile);
                        for (int s = 0; s < kSlabs; ++s)
                            tma_load_4d(sQ + t * L::kTile + s * L::kSlab, &tmQ, bar_q_full + 8 * (t), s * 64, w.bidh,
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
              __syncwarp();
              if (pass == 0) { pass_boundary(); if (pass_count(1) == 0) break; }
            }
        } else if (warp == 16 || warp == kMma1Warp) {
            // ===================== MMA issuers: warp 16 -> tile 0, warp 17 -> tile 1 =====================
            // The tensor pipe work of a key step is two BLOCKS, (tile 0: four P quarters + next S) (tile 1: ...), strictly
            // alternating: that order keeps the two tiles half a period apart — one in its softmax while the other's MMAs
            // run.  With ONE issuing warp (round 2, first half) clock64 showed 350-550 cycles between the last MMA of a block
            // and the first of the next: issue blocks while more than 2-3 MMAs are queued, so the warp comes back with
            // < 200 cycles of work in the pipe and then has three commits, a V wait, the P probe and the descriptor
            // arithmetic to get through — the pipe ran dry once per block (period 2 x (1024 + 350), tensor pipe 65 % busy).
            // Now every tile has its own issuing warp that does all its waiting IN ADVANCE and then waits for a token the
            // other warp passes on right behind its last MMA (mbarrier arrive -> try_wait wake-up).  Without the token
            // (tried in round 2, profiles/r02_run5.log) the tiles drift into the same phase: 3760 cycles per period.
            // Block order inside an item: (0,0) (1,0) (0,1) (1,1) ...; tile t has nb_t blocks, so
            //   (0,j) waits for (1,j-1) iff j >= 1 and j-1 < nb1, and wakes (1,j) iff j < nb1;
            //   (1,j) waits for (0,j) iff j < nb0, and wakes (0,j+1) iff j+1 < nb0;
            // the two warps meet on a named barrier at the end of every item (a token may never be passed twice before it
            // has been picked up: the mbarrier phase parity would alias).
            const int t = (warp == 16) ? 0 : 1;
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc(kBf16, kBlockM, kBlockN, false, false);
            constexpr uint32_t idesc_pv = make_idesc(kBf16, kBlockM, D, false, true);
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t q_lo = __shfl_sync(0xffffffffu, desc_lo(sQ, 16), 0);
            const uint32_t kv_lo = __shfl_sync(0xffffffffu, desc_lo(sKV, 16), 0);
            const uint32_t v_lo = __shfl_sync(0xffffffffu, desc_lo(sKV, L::kSlab), 0);
            constexpr uint32_t kTile16 = L::kTile >> 4;
            int kv_i = 0;                 // ring index of K_0 of the current item
            int its = 0;                  // S_t / P_t steps so far (barrier parities)
            int nitem = 0;                // items in which this tile had keys (Q / O barrier parities)
            int ntok = 0;                 // tokens picked up so far
            for (int pass = 0; pass < 2; ++pass) {
              const int cnt = pass_count(pass);
              for (int idx = 0; idx < cnt; ++idx) {
                const int n = pass_item(pass, idx);
                const WorkItem w = decode_item(ts, n, p.h, p.is_causal != 0);
                const ItemGeom g = item_geom(p, w);
                const int nb0 = __shfl_sync(0xffffffffu, g.nblk[0], 0);
                const int nb1 = __shfl_sync(0xffffffffu, g.nblk[1], 0);
                const int nbmax = max(nb0, nb1);
                if (__shfl_sync(0xffffffffu, (int)g.skip, 0) || nbmax == 0) continue;
                const int nbt = t == 0 ? nb0 : nb1;
                auto kv_slot = [&](int i) { return (kv_i + i) % kStages; };
                auto wait_kv = [&](int i) { mbar_wait(bar_kv_full + 8 * (kv_slot(i)), (((kv_i + i) / kStages) & 1)); };
                auto commit = [&](uint32_t bar) { if (leader) tc_commit(bar); };
                // S_t = Q_t K_j^T, then the token.  Passing it a few MMAs BEFORE the end of the block (so that the other warp's
                // first MMAs queue up right behind this block's last ones) measured much slower — C2 1121 instead of 1356
                // TFLOP/s, profiles/r02_run17.log: with MMAs of two warps in flight at the same time every MMA took ~145
                // cycles instead of 64 (clock64; the pipe seems to alternate between the issuing warps and drain in between).
                auto issue_s = [&](int j, bool give) {
                    if (leader) {
                        const uint32_t qa = q_lo + t * kTile16;
                        const uint32_t ka = kv_lo + kv_slot(2 * j) * kTile16;
#pragma unroll
                        for (int kk = 0; kk < D / 16; ++kk) {
                            const uint32_t off = ((kk >> 2) * L::kSlab + (kk & 3) * 32) >> 4;
                            umma_ss(tm + kTmemS0 + t * 128, desc_make(qa + off, kDescHiK), desc_make(ka + off, kDescHiK),
                                    idesc_s, kk > 0);
                        }
                    }
                    if (leader) {
                        if (give) mbar_arrive(bar_tok + 8 * (t ^ 1));
                        tc_commit(bar_s_full + 8 * (t));
                    }
                };
                // O_t += P_t[:, quarter q] V_j[quarter q]: key rows [16q, 16q+16) (P written by half 0) and
                // [64+16q, 64+16q+16) (half 1); P quarter (hh, q) sits at S_t + 64 hh + 8 q
                auto issue_pv = [&](int j, int q) {
                    if (leader) {
                        const uint32_t va = v_lo + kv_slot(2 * j + 1) * kTile16;
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh)
                            umma_ts(tm + kTmemO0 + t * 128, tm + kTmemS0 + t * 128 + hh * 64 + q * 8,
                                    desc_make(va + (hh * 4 + q) * (2048 >> 4), kDescHiK), idesc_pv, (j > 0 || q > 0 || hh > 0));
                    }
                };
                if (nbt > 0) {
                    wait_kv(0);
                    mbar_wait(bar_q_full + 8 * (t), nitem & 1);
                    tc_fence_after();
                    issue_s(0, false);
                    if (nbt == 1) commit(bar_q_empty + 8 * (t));          // Q_t may be overwritten by the next item
                    commit(bar_kv_empty + 8 * (kv_slot(0)));
                }
                for (int j = 0; j < nbt; ++j) {
                    // everything this block needs, before asking for the token
                    wait_kv(2 * j + 1);                                             // V_j
                    if (j + 1 < nbt) wait_kv(2 * j + 2);                            // K_j+1
                    if (j == 0) mbar_wait(bar_o_empty + 8 * (t), (nitem & 1) ^ 1);        // O_t of the previous item has been read out
                    const bool take = (t == 0) ? (j >= 1 && j - 1 < nb1) : (j < nb0);
                    if (take) { mbar_wait(bar_tok + 8 * (t), ntok & 1); ++ntok; }
                    // The softmax warps hand P over quarter by quarter, in order.  When all four are there one probe of the
                    // last quarter replaces four waits.  (Issuing "as many quarters as are ready" back to back, or the first
                    // quarter before the token, measured 7-9 % SLOWER, profiles/r02_run18.log: P V MMAs that run underneath
                    // the tile's own remaining exponentials compete with its TMEM loads and stores.  Handing P over only
                    // once per key step: 1-5 % slower, r02_run19.log.  The token through a hardware named barrier instead of an
                    // mbarrier, with or without the first quarter waited for ahead of it: -1.2 % ... +0.5 %, r02_run21.log.)
                    const bool give = (t == 0) ? (j < nb1) : (j + 1 < nb0);
                    const bool last = j + 1 == nbt;
                    if (mbar_test_wait(bar_p + 8 * (4 * t + 3), (its + j) & 1)) {
                        tc_fence_after();
                        if (lane == 0) FA_TRACE_EVENT(4 + t, its + j, 3);
#pragma unroll
                        for (int q = 0; q < 4; ++q) issue_pv(j, q);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            mbar_wait(bar_p + 8 * (4 * t + q), (its + j) & 1);
                            tc_fence_after();
                            if (lane == 0) FA_TRACE_EVENT(4 + t, its + j, q);      // rows 4 / 5: MMA warp of tile 0 / 1
                            issue_pv(j, q);
                        }
                    }
                    if (last && give && leader) mbar_arrive(bar_tok + 8 * (t ^ 1));       // last block of the tile: no S follows
                    if (lane == 0) FA_TRACE_EVENT(4 + t, its + j, 4);
                    if (!last) {
                        issue_s(j + 1, give);
                    } else {
                        commit(bar_o_full + 8 * (t));
                    }
                    if (lane == 0) FA_TRACE_EVENT(4 + t, its + j, 6);
                    if (j + 2 == nbt) commit(bar_q_empty + 8 * (t));                      // S_t of the last step issued
                    commit(bar_kv_empty + 8 * (kv_slot(2 * j + 1)));
                    if (j + 1 < nbt) commit(bar_kv_empty + 8 * (kv_slot(2 * j + 2)));
                    __syncwarp();
                }
                // K/V tiles only the other query tile needs (causal diagonal, ragged tails): let go of them in ring order,
                // each only once it has landed — its slot's previous release is complete by then
                for (int j = nbt; j < nbmax; ++j) {
                    wait_kv(2 * j);
                    if (leader) mbar_arrive(bar_kv_empty + 8 * (kv_slot(2 * j)));
                    wait_kv(2 * j + 1);
                    if (leader) mbar_arrive(bar_kv_empty + 8 * (kv_slot(2 * j + 1)));
                    __syncwarp();
                }
                kv_i += 2 * nbmax;
                its += nbt;
                nitem += (nbt > 0);
                named_bar_sync(13, 64);                                             // the two issuing warps leave an item together
              }
              if (pass == 0) { pass_boundary(); if (pass_count(1) == 0) break; }
            }
        } else {
            // ===================== store warp(s): TMA store of the staged O tiles =====================
            // The softmax warpgroups only write the staging tile and arrive on a named barrier; issuing the bulk store and
            // waiting for the engine to read 32 KB of shared memory (~1800 cycles) is this warp's job.  When a softmax
            // thread did it, its whole warp sat in that wait and — every P quarter needs all eight warps of a tile — held
            // up the tile's first key steps of the next item (clock64: the next item's first S was picked up ~1200 cycles
            // after the epilogue had finished, profiles/r02_run3.log).
            // One warp serves both tiles: whoever holds the staging lock describes the tile in store_desc and arrives on
            // barrier 11; this warp only has to know HOW MANY whole tiles the CTA's items produce in this pass.
            for (int pass = 0; pass < 2; ++pass) {
              const int cnt = pass_count(pass);
              int n_store = 0;
              for (int idx = 0; idx < cnt; ++idx) {
                const int n = pass_item(pass, idx);
                const WorkItem w = decode_item(ts, n, p.h, p.is_causal != 0);
                const ItemGeom g = item_geom(p, w);
                if (g.skip) continue;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int mt = g.m0 + t * kBlockM;
                    if (mt >= g.sq_b || g.nblk[t] == 0) continue;
                    n_store += ((mt + kBlockM <= g.sq_b) || (p.cu_q == nullptr)) ? 1 : 0;   // ragged varlen tails are stored by the softmax threads
                }
              }
#pragma unroll 1
              for (int i = 0; i < n_store; ++i) {
                named_bar_sync(11, 2 * kBlockM + 32);              // staging tile written and fenced by a tile's 256 threads
                if (lane == 0) {
                    const int bidh = lds32(store_desc), row0 = lds32(store_desc + 4), tb = lds32(store_desc + 8);
#pragma unroll
                    for (int sl = 0; sl < kSlabs; ++sl)
                        tma_store_4d(&tmO, sStage + sl * L::kSlab, sl * 64, bidh, row0, tb);
                    tma_store_commit();
                    tma_store_wait_read<0>();                      // staging tile has been read; global writes complete later
                    __threadfence_block();
                    atoms_exch(stage_lock, 0);                     // the other tile's epilogue (or this tile's next one) may take it
                }
                __syncwarp();
              }
              if (lane == 0) tma_store_wait<0>();                  // all bulk stores have landed before a retry rewrites the tile / the CTA retires
              __syncwarp();
              if (pass == 0) { pass_boundary(); if (pass_count(1) == 0) break; }
            }
        }
    } else {
        // ========== softmax warpgroups: warpgroup 2t+hh owns columns [64hh, 64hh+64) of tile slot t, one thread per row ==========
        // Register budget: five warps per SM sub-partition launch with 96 registers each; setmaxnreg.inc can only draw what
        // setmaxnreg.dec released (the launch-time slack of the register file is NOT in the pool: a 112/64 split deadlocks,
        // profiles/r01s2_run1_setmaxnreg_hang.log), so warpgroup 4 gives up 32 per thread and each softmax warp takes 8.
        setmaxnreg_inc<104>();
        const int t = wg >> 1;
        const int hh = wg & 1;
        const int wq = warp & 3;                         // TMEM lane quadrant = SM sub-partition
        const int r_in_tile = tid & 127;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        const uint32_t tS = tmem_base + lane_base + kTmemS0 + t * 128 + hh * 64;       // own scores; own P goes to the same place
        const uint32_t tO = tmem_base + lane_base + kTmemO0 + t * 128 + hh * kHalfD;   // own half of the O row
        const uint32_t x_own = xch + ((t * 2 + hh) * kBlockM + r_in_tile) * 4;
        const uint32_t pair_bar = 1 + t * 4 + wq;       // named barriers 1..8: the two warps that share 32 rows (64 threads)
        const uint32_t tile_bar = 9 + t;                // named barriers 9, 10: the two warpgroups of a tile (256 threads); 11, 12: + its store warp
        uint16_t* o_base = reinterpret_cast<uint16_t*>(p.o);
        const float c2 = p.scale_log2;
        int its = 0;       // S_t steps so far
        int nitem = 0;     // items with keys finished by this slot
        uint32_t hs16 = 0; // own sum of the previous key step, top 16 bits (what the peer thread reads one step late)

        // P = 2^(s*c2 + neg) for one 16-column chunk; kEmu selects how many of every 8 column pairs go through the
        // Cody-Waite + degree-3 polynomial path on the FMA pipe instead of MUFU.EX2 (|rel err| < 7.5e-5, far below the
        // 16-bit rounding of P).  The argument never exceeds 8 (exact max, lazy reference), so only the lower end is
        // clamped (-inf for masked columns; below -125 the exponent add would leave the normal range).
        auto exp_half = [&](auto half_tag, auto spec_tag, const float (&s)[16], const float neg, uint32_t (&pk8)[8], float2& sum) {
            constexpr int kHalf = decltype(half_tag)::value;         // pairs [4 kHalf, 4 kHalf + 4) of the chunk's 8
            constexpr bool kSpec = decltype(spec_tag)::value;        // speculative step: the argument is not bounded above
            const float2 c2v = make_float2(c2, c2);
            const float2 negv = make_float2(neg, neg);
            // kEmu -> emulated pairs of every 8 (spread evenly): 4 -> {0}, 1 -> {0,4}, 3 -> {0,3,6}, 2 -> {0,2,4,6}
            constexpr int kEmu8 = kEmu == 1 ? 2 : kEmu == 2 ? 4 : kEmu == 3 ? 3 : kEmu == 4 ? 1 : 0;
#pragma unroll
            for (int i = 4 * kHalf; i < 4 * kHalf + 4; ++i) {
                float2 x = __ffma2_rn(make_float2(s[2 * i], s[2 * i + 1]), c2v, negv);
                float2 pp;
                if (((i * kEmu8) & 7) < kEmu8) {
                    const float2 magic = make_float2(12582912.f, 12582912.f);       // 1.5 * 2^23
                    x = make_float2(fmaxf(x.x, -125.f), fmaxf(x.y, -125.f));
                    // above 126 the exponent add below would wrap around (2^141 came out as -2^-116 in round 1): clamped,
                    // the value is ~2^126 and trips the overflow check of the speculative step like MUFU's +inf does
                    if constexpr (kSpec) x = make_float2(fminf(x.x, 126.f), fminf(x.y, 126.f));
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
                    pp = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                }
                sum = __fadd2_rn(sum, pp);
                pk8[i] = pack2<kBf16>(pp.x, pp.y);
            }
        };
        using h0 = std::integral_constant<int, 0>;
        using h1 = std::integral_constant<int, 1>;
        auto mask_chunk = [&](float (&s)[16], const int lim_c) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c > lim_c) s[c] = -INFINITY;
        };
        auto max_chunk = [&](const float (&s)[16]) -> float {
            float ma = fmaxf(s[0], s[1]), mb = fmaxf(s[2], s[3]);
#pragma unroll
            for (int c = 4; c < 16; c += 4) {
                ma = fmaxf(ma, fmaxf(s[c], s[c + 1]));
                mb = fmaxf(mb, fmaxf(s[c + 2], s[c + 3]));
            }
            return fmaxf(ma, mb);
        };
        // clock64 builds: rows 0 / 1 = warp 0 of tile 0 / 1 (column half 0, lane quadrant 0), rows 2 / 3 = warp 7 of the tile
        // (column half 1, quadrant 3); events 0 S full, 1 max exchanged, 2..5 quarter q handed over, 6 O full, 7 epilogue done
        const int trole = (hh == 0 && wq == 0) ? t : (hh == 1 && wq == 3) ? 2 + t : 99;
        int tr_step = 0;
        const int erole = (hh == 0 && wq == 0) ? 6 + t : 99;   // rows 6 / 7: epilogue stamps of warp 0 of tile 0 / 1, one row per item
        (void)trole; (void)tr_step; (void)erole;
        // own sum of step j, truncated to its top 16 bits, into half (j & 1) of the own exchange slot
        auto publish_sum = [&](const int j, const float2& sum) {
            hs16 = __float_as_uint(sum.x + sum.y) >> 16;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(x_own + ((j & 1) << 1)), "h"((uint16_t)hs16) : "memory");
        };
        constexpr float kOverflowAt = kBf16 ? 1.2676506e30f /* 2^100 */ : 32768.f /* fp16 P <= 65504 */;
        // A quarter of P leaves in two steps: the TMEM store is issued as soon as the chunk's exponentials are done, the
        // hand-over to the MMA warp (wait::st, fence, one elected arrive per warp) half a chunk later, underneath the next
        // chunk's exponentials.  Waiting for the store right behind its issue cost ~130 cycles per quarter with both warps
        // of a sub-partition stalled at the same time and the MUFU pipe idle (clock64: pass 2 took 1520 cycles instead of
        // ~1000, profiles/r02_run1.log).
        auto arrive_quarter = [&](const int q) {
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p + 8 * (4 * t + q));
            if (lane == 0) FA_TRACE_EVENT(trole, tr_step, 2 + q);
        };

        for (int pass = 0; pass < 2; ++pass) {
          const bool safe = pass == 1 || exact0;     // exact steps only (pass 1 redoes the items whose speculation overflowed)
          const int cnt = pass_count(pass);
          for (int idx = 0; idx < cnt; ++idx) {
            const int n = pass_item(pass, idx);
            // Only what the key loop needs stays live across it (n_t, the first step that needs a mask, the thread's column
            // limit): with 104 registers and no L1 (shared memory takes all of it) every spilled value is an L2 round trip.
            // The epilogue decodes the item again.
            int n_t, j_mask, lim0;
            {
                const WorkItem w = decode_item(ts, n, p.h, p.is_causal != 0);
                const ItemGeom g = item_geom(p, w);
                if (g.skip) continue;
                const int mt = g.m0 + t * kBlockM;
                if (mt >= g.sq_b) continue;                  // this slot has no rows in this item (nblk[t] == 0 too)
                n_t = t == 0 ? g.nblk[0] : g.nblk[1];            // (a run-time index would put the struct into local memory)
                const int row = mt + r_in_tile;
                if (n_t == 0) {
                    // rows exist but see no key: O = 0, LSE = 0 (each half clears its D/2 columns)
                    if (row < g.sq_b) {
                        const int64_t o_row_base = (p.cu_q != nullptr) ? (int64_t)g.q_row0 : (int64_t)w.bidb * p.sq;
                        uint16_t* o_row = o_base + ((o_row_base + row) * p.h + w.bidh) * D + hh * kHalfD;
#pragma unroll
                        for (int ch = 0; ch < kHalfD / 8; ++ch) *(reinterpret_cast<uint4*>(o_row) + ch) = make_uint4(0, 0, 0, 0);
                        if (hh == 0) p.lse[((int64_t)w.bidb * p.h + w.bidh) * p.sq + row] = 0.f;
                    }
                    continue;
                }
                int col_limit = g.sk_b - 1;                  // last visible key of this row
                if (p.is_causal) col_limit = min(col_limit, row + g.causal_off);
                lim0 = col_limit - hh * 64;                  // ... relative to this thread's first column of key tile 0
                // key tile j needs a mask iff 128 j + 128 > sk_b (tail) or, causal, 128 j + 127 > mt + causal_off (diagonal)
                int first_masked = g.sk_b / kBlockN;         // floor: tiles below are entirely inside the sequence
                if (p.is_causal) first_masked = min(first_masked, (mt + g.causal_off + 1) / kBlockN);
                j_mask = max(first_masked, 0);
            }
            // softmax state of the row: reference exponent neg = -m_ref * c2 (identical in both threads of the row), own partial sum
            float neg = 0.f, l_run = 0.f;
            bool has_ref = false;
            bool overflowed = false;     // some speculative step of this thread overflowed in this item
            hs16 = 0;

            // ------------------------------------------------------------------------------------------------------
            // exact step: row max first, then the exponentials.  Every step of the retry pass (pass 1), and the first
            // step of every item.
            // ------------------------------------------------------------------------------------------------------
            auto key_step = [&](auto mask_tag, const int j) {
                constexpr bool need_mask = decltype(mask_tag)::value;   // masked and unmasked steps get separate straight-line copies
                using spec = std::false_type;
                tr_step = its + j;
                const int lim = lim0 - j * kBlockN;            // last visible column of this thread's 64 (may be < 0 or >= 64)
                if (j == 0 && (tid & 31) == 0) FA_TRACE_EVENT(trole, its, 7);   // first step of an item: about to wait for its S
                mbar_wait(bar_s_full + 8 * (t), (its + j) & 1);
                tc_fence_after();
                if ((tid & 31) == 0) FA_TRACE_EVENT(trole, its + j, 0);

                // ---- pass 1: exact row max of the 64 own scores; chunk 0 stays in registers for pass 2 ----
                float sa[16];
                float mx;
                {
#if FA_P4_PASS1_WIDE
                    float sc[16], sd[16], sb[16];
                    tmem_ld16(tS + 32, *reinterpret_cast<uint32_t(*)[16]>(&sc[0]));
                    tmem_ld16(tS + 48, *reinterpret_cast<uint32_t(*)[16]>(&sd[0]));
                    tmem_ld16(tS + 16, *reinterpret_cast<uint32_t(*)[16]>(&sb[0]));
                    tmem_ld16(tS, *reinterpret_cast<uint32_t(*)[16]>(&sa[0]));
                    tmem_wait_ld();
                    if constexpr (need_mask) { mask_chunk(sc, lim - 32); mask_chunk(sd, lim - 48); mask_chunk(sb, lim - 16); mask_chunk(sa, lim); }
                    mx = fmaxf(fmaxf(max_chunk(sc), max_chunk(sd)), fmaxf(max_chunk(sb), max_chunk(sa)));
#else
                    float sc[16], sd[16], sb[16];
                    tmem_ld16(tS + 32, *reinterpret_cast<uint32_t(*)[16]>(&sc[0]));
                    tmem_ld16(tS + 48, *reinterpret_cast<uint32_t(*)[16]>(&sd[0]));
                    tmem_wait_ld();
                    tmem_ld16(tS + 16, *reinterpret_cast<uint32_t(*)[16]>(&sb[0]));
                    if constexpr (need_mask) { mask_chunk(sc, lim - 32); mask_chunk(sd, lim - 48); }
                    mx = fmaxf(max_chunk(sc), max_chunk(sd));
                    tmem_wait_ld();
                    tmem_ld16(tS, *reinterpret_cast<uint32_t(*)[16]>(&sa[0]));
                    if constexpr (need_mask) mask_chunk(sb, lim - 16);
                    mx = fmaxf(mx, max_chunk(sb));
                    tmem_wait_ld();
                    if constexpr (need_mask) mask_chunk(sa, lim);
                    mx = fmaxf(mx, max_chunk(sa));
#endif
                }
                {
                    sts32f(x_own, mx);
                    named_bar_sync(pair_bar, 64);
                    float mp;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mp) : "r"(x_own ^ (kBlockM * 4)));
                    mx = fmaxf(mx, mp);
                    // the 16-bit halves of this slot carry the speculative steps' sums: nobody may publish into it before
                    // both threads of the row have read the maxima
                    named_bar_sync(pair_bar, 64);
                }
                if ((tid & 31) == 0) FA_TRACE_EVENT(trole, its + j, 1);
                // lazy reference: it moves only when the row max is more than 2^8 above it (always on the first visible key)
                const float xm = fmaf(mx, c2, neg);
                const bool need = has_ref ? (xm > kRescaleThreshold) : (mx > -INFINITY);
                if (__any_sync(0xffffffffu, need)) {
                    // both warps of the pair take this path together (they see the same 32 row maxima)
                    float alpha = 1.f;
                    if (need) {
                        alpha = has_ref ? fast_exp2(-xm) : 0.f;      // 2^((m_old - m_new) c2); nothing accumulated yet without a reference
                        neg = -mx * c2;
                        has_ref = true;
                        l_run *= alpha;
                    }
                    if (j > 0) {
#pragma unroll 1
                        for (int c = 0; c < kHalfD / 16; ++c) {
                            uint32_t o[16];
                            tmem_ld16(tO + c * 16, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st16(tO + c * 16, o);
                        }
                        tmem_wait_st();
                        // P_t V of this step accumulates into ALL of O_t: neither half may release any P before both
                        // halves of the row block have finished rescaling
                        tc_fence_before();
                        named_bar_sync(pair_bar, 64);
                        tc_fence_after();
                    }
                }

                // ---- pass 2: exponentials; the next chunk's scores are in flight, the previous quarter's hand-over is folded in ----
                float2 sum = make_float2(0.f, 0.f);
                {
                    float sb[16];
                    uint32_t pka[8], pkb[8];
                    tmem_ld16(tS + 16, *reinterpret_cast<uint32_t(*)[16]>(&sb[0]));
                    exp_half(h0{}, spec{}, sa, neg, pka, sum);
                    exp_half(h1{}, spec{}, sa, neg, pka, sum);
                    tmem_st8(tS, pka);
                    tmem_wait_ld();
                    tmem_ld16(tS + 32, *reinterpret_cast<uint32_t(*)[16]>(&sa[0]));
                    if constexpr (need_mask) mask_chunk(sb, lim - 16);
                    exp_half(h0{}, spec{}, sb, neg, pkb, sum);
                    arrive_quarter(0);
                    exp_half(h1{}, spec{}, sb, neg, pkb, sum);
                    tmem_st8(tS + 8, pkb);
                    tmem_wait_ld();
                    tmem_ld16(tS + 48, *reinterpret_cast<uint32_t(*)[16]>(&sb[0]));
                    if constexpr (need_mask) mask_chunk(sa, lim - 32);
                    exp_half(h0{}, spec{}, sa, neg, pka, sum);
                    arrive_quarter(1);
                    exp_half(h1{}, spec{}, sa, neg, pka, sum);
                    tmem_st8(tS + 16, pka);
                    tmem_wait_ld();
                    if constexpr (need_mask) mask_chunk(sb, lim - 48);
                    exp_half(h0{}, spec{}, sb, neg, pkb, sum);
                    arrive_quarter(2);
                    exp_half(h1{}, spec{}, sb, neg, pkb, sum);
                    publish_sum(j, sum);                  // before the last hand-over: the peer reads it after the next S is full
                    tmem_st8(tS + 24, pkb);
                    arrive_quarter(3);
                }
                l_run += sum.x + sum.y;
            };

            // ------------------------------------------------------------------------------------------------------
            // speculative step ("late-agreed reference"): no pass 1, no exchange on the critical path.  The exponentials
            // use the reference agreed so far; the two threads of a row publish their step sums and read each other's ONE
            // STEP LATE: if the row's sum of step j-1 exceeded 2^9 both shift the reference up by floor(log2 sum) (identical
            // arithmetic on identical inputs -> identical reference, no vote) and rescale their half of O and their partial
            // l — an exact power of two.  What cannot be fixed one step late is an overflow inside a single key tile
            // (bf16: a row sum above 2^100, fp16: above 2^15, i.e. scores jumping by that much within 128 keys): the thread
            // raises the CTA's retry flag (after the key loop), its rows come out as garbage, and pass 1 redoes the CTA's items
            // with exact steps.
            // ------------------------------------------------------------------------------------------------------
            auto spec_step = [&](auto mask_tag, const int j) {
                constexpr bool need_mask = decltype(mask_tag)::value;
                using spec = std::true_type;
                tr_step = its + j;
                const int lim = lim0 - j * kBlockN;
                mbar_wait(bar_s_full + 8 * (t), (its + j) & 1);
                tc_fence_after();
                if ((tid & 31) == 0) FA_TRACE_EVENT(trole, its + j, 0);
                float sa[16], sb[16];
                tmem_ld16(tS, *reinterpret_cast<uint32_t(*)[16]>(&sa[0]));
                tmem_ld16(tS + 16, *reinterpret_cast<uint32_t(*)[16]>(&sb[0]));
                {   // late agreement on the reference (underneath the TMEM loads)
                    uint32_t peer16;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(peer16) : "r"((x_own ^ (kBlockM * 4)) + (((j - 1) & 1) << 1)));
                    const float tot = __uint_as_float(hs16 << 16) + __uint_as_float(peer16 << 16);
                    const bool up = tot > 512.f;
                    if (__any_sync(0xffffffffu, up)) {
                        int e = 0;
                        if (up) e = min((int)((__float_as_uint(tot) >> 23) & 0xffu) - 127, 120);   // floor(log2 tot); inf / garbage: the row is flagged anyway
                        const float sc = __uint_as_float((uint32_t)(127 - e) << 23);              // 2^-e
                        neg -= (float)e;
                        l_run *= sc;
                        tmem_wait_ld();
#pragma unroll 1
                        for (int c = 0; c < kHalfD / 16; ++c) {
                            uint32_t o[16];
                            tmem_ld16(tO + c * 16, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
                            tmem_st16(tO + c * 16, o);
                        }
                        tmem_wait_st();
                        tc_fence_before();
                        named_bar_sync(pair_bar, 64);      // P V of this step accumulates into all of O_t: both halves rescaled first
                        tc_fence_after();
                    }
                }
                if ((tid & 31) == 0) FA_TRACE_EVENT(trole, its + j, 1);
                float2 sum = make_float2(0.f, 0.f);
                {
                    uint32_t pka[8], pkb[8];
                    tmem_wait_ld();
                    if constexpr (need_mask) mask_chunk(sa, lim);
                    exp_half(h0{}, spec{}, sa, neg, pka, sum);
                    exp_half(h1{}, spec{}, sa, neg, pka, sum);
                    tmem_st8(tS, pka);
                    tmem_ld16(tS + 32, *reinterpret_cast<uint32_t(*)[16]>(&sa[0]));
                    if constexpr (need_mask) mask_chunk(sb, lim - 16);
                    exp_half(h0{}, spec{}, sb, neg, pkb, sum);
                    arrive_quarter(0);
                    exp_half(h1{}, spec{}, sb, neg, pkb, sum);
                    tmem_st8(tS + 8, pkb);
                    tmem_wait_ld();
                    tmem_ld16(tS + 48, *reinterpret_cast<uint32_t(*)[16]>(&sb[0]));
                    if constexpr (need_mask) mask_chunk(sa, lim - 32);
                    exp_half(h0{}, spec{}, sa, neg, pka, sum);
                    arrive_quarter(1);
                    exp_half(h1{}, spec{}, sa, neg, pka, sum);
                    tmem_st8(tS + 16, pka);
                    tmem_wait_ld();
                    if constexpr (need_mask) mask_chunk(sb, lim - 48);
                    exp_half(h0{}, spec{}, sb, neg, pkb, sum);
                    arrive_quarter(2);
                    exp_half(h1{}, spec{}, sb, neg, pkb, sum);
                    publish_sum(j, sum);
                    tmem_st8(tS + 24, pkb);
                    arrive_quarter(3);
                }
                const float hs = sum.x + sum.y;
                // overflow (also inf / NaN), or no reference yet (only an exact step can establish one; with prefix masks a row
                // that saw no key in the first tile sees none at all and never gets here with more than one key step)
                overflowed = overflowed || !(hs <= kOverflowAt) || (!has_ref && hs != 0.f);
                l_run += hs;
            };
            for (int j = 0; j < n_t; ++j) {
                if (safe || j == 0) {
                    if (j >= j_mask) key_step(std::true_type{}, j);
                    else key_step(std::false_type{}, j);
                } else {
                    if (j >= j_mask) spec_step(std::true_type{}, j);
                    else spec_step(std::false_type{}, j);
                }
            }
            its += n_t;
            if (overflowed) sts32(retry_flag, 1);            // pass 1 redoes the CTA's items with exact steps

            // ---- epilogue: O_t / l -> 16 bit -> staging tile (128B-swizzled, the TMA layout) -> TMA store by warp 19 ----
            // Measured alternatives (profiles/r02_run24.log, r02_run25.log): storing O straight from registers with 32-byte
            // global stores (no staging tile, no lock, no proxy fence) — the scattered stores back up in the LSU for ~2000
            // cycles per tile: neutral at seqlen >= 4096, 2.5-4 % slower at seqlen 512-1024; looking up the NEXT item before
            // this epilogue (the work-list decode is a ~1000-cycle dependent integer chain per item) — +0.3-0.9 % at head_dim
            // 128, -0.7 ... -2.7 % at head_dim 64 (two more values spilled), not kept.
            if ((tid & 31) == 0) FA_TRACE_EVENT(erole, nitem, 0);     // rows 6 / 7: epilogue of item `nitem` (0 start, 1 O full, 2 l exchanged,
            int n_again = n;                                           //   3 staging tile free, 4 O in registers, 5 staged, 6 handed over)
            asm volatile("" : "+r"(n_again));                  // opaque copy: keeps the geometry from living across the key loop
            const WorkItem w = decode_item(ts, n_again, p.h, p.is_causal != 0);
            const ItemGeom g = item_geom(p, w);
            const int mt = g.m0 + t * kBlockM;
            const int row = mt + r_in_tile;
            const int64_t o_row_base = (p.cu_q != nullptr) ? (int64_t)g.q_row0 : (int64_t)w.bidb * p.sq;
            float* lse_row = p.lse + ((int64_t)w.bidb * p.h + w.bidh) * p.sq;
            sts32f(x_own, l_run);
            named_bar_sync(pair_bar, 64);
            float l_peer;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l_peer) : "r"(x_own ^ (kBlockM * 4)));
            const float l_tot = l_run + l_peer;
            const bool row_empty = !has_ref || !(l_tot > 0.f);   // no visible key: O = 0, LSE = 0
            const float inv_l = row_empty ? 0.f : (1.f / l_tot);
            if ((tid & 31) == 0) FA_TRACE_EVENT(erole, nitem, 2);
            mbar_wait(bar_o_full + 8 * (t), nitem & 1);
            tc_fence_after();
            if ((tid & 31) == 0) FA_TRACE_EVENT(erole, nitem, 1);
            if (hh == 0 && wq == 0) {    // take the staging tile (the other tile's store may still be reading it); the whole warp
                int got;                 // spins together: bar.sync below is warp-aligned
                do {
                    got = 0;
                    if (lane == 0) got = (atoms_cas(stage_lock, 0, 1) == 0);
                    got = __shfl_sync(0xffffffffu, got, 0);
                    if (!got) __nanosleep(32);
                } while (!got);
            }
            named_bar_sync(tile_bar, 2 * kBlockM);
            if ((tid & 31) == 0) FA_TRACE_EVENT(erole, nitem, 3);
            const uint32_t stage = sStage;
#pragma unroll
            for (int c = 0; c < kHalfD / 32; ++c) {
                uint32_t o[32];
                tmem_ld32(tO + c * 32, o);
                tmem_wait_ld();
                if (c == kHalfD / 32 - 1) {     // O_t is in registers: the next item's first P V may overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_o_empty + 8 * (t));
                    if (lane == 0) FA_TRACE_EVENT(erole, nitem, 4);
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 v;
                    v.x = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 0]) * inv_l, __uint_as_float(o[q4 * 8 + 1]) * inv_l);
                    v.y = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 2]) * inv_l, __uint_as_float(o[q4 * 8 + 3]) * inv_l);
                    v.z = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 4]) * inv_l, __uint_as_float(o[q4 * 8 + 5]) * inv_l);
                    v.w = pack2<kBf16>(__uint_as_float(o[q4 * 8 + 6]) * inv_l, __uint_as_float(o[q4 * 8 + 7]) * inv_l);
                    const int chunk = (hh * kHalfD) / 8 + c * 4 + q4;     // 16-byte chunk of the output row; 8 per 64-column slab
                    sts128u(stage + (chunk >> 3) * L::kSlab + r_in_tile * 128 + (((chunk & 7) ^ (r_in_tile & 7)) << 4), v);
                }
            }
            if (hh == 0 && row < g.sq_b) lse_row[row] = row_empty ? 0.f : fmaf(-neg, 0.6931471805599453f, logf(l_tot));   // m_ref / sqrt(d) + ln l
            if ((tid & 31) == 0) FA_TRACE_EVENT(erole, nitem, 5);
            if (hh == 0 && r_in_tile == 0) {                   // tell the store warp what the staging tile holds
                sts32(store_desc, w.bidh); sts32(store_desc + 4, g.q_row0 + mt); sts32(store_desc + 8, g.tma_b);
                __threadfence_block();
            }
            fence_proxy_async_smem();                          // generic-proxy writes -> visible to the TMA engine
            const bool whole_tile = (mt + kBlockM <= g.sq_b) || (p.cu_q == nullptr);   // dense: TMA clips rows >= seqlen_q itself
            if (whole_tile) {
                // hand the tile to the store warp and move on: it checks the overflow flag, issues the bulk store and
                // releases the staging tile once the engine has read it
                named_bar_arrive(11, 2 * kBlockM + 32);
            } else {
                // ragged varlen tail: a TMA box would spill into the next sequence -> predicated coalesced stores
                constexpr int kChunksPerRow = D / 8;
                named_bar_sync(tile_bar, 2 * kBlockM);
                for (int idx2 = hh * kBlockM + r_in_tile; idx2 < kBlockM * kChunksPerRow; idx2 += 2 * kBlockM) {
                    const int rr = idx2 / kChunksPerRow, ch = idx2 % kChunksPerRow;
                    if (mt + rr < g.sq_b) {
                        const uint4 v = lds128u(stage + (ch >> 3) * L::kSlab + rr * 128 + (((ch & 7) ^ (rr & 7)) << 4));
                        *(reinterpret_cast<uint4*>(o_base + ((o_row_base + mt + rr) * p.h + w.bidh) * D) + ch) = v;
                    }
                }
                named_bar_sync(tile_bar, 2 * kBlockM);
                if (hh == 0 && r_in_tile == 0) {
                    atoms_exch(stage_lock, 0);
                }
                __syncwarp();
            }
            if ((tid & 31) == 0) FA_TRACE_EVENT(trole, its - 1, 6);
            if ((tid & 31) == 0) FA_TRACE_EVENT(erole, nitem, 6);
            ++nitem;
          }
          if (pass == 0) { pass_boundary(); if (pass_count(1) == 0) break; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int D, bool kBf16, int kEmu>
int launch_fwd_p4(const fa_fwd_params* p, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, FwdParams kp,
                  cudaStream_t stream) {
    using L = P4Smem<D>;
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
    auto kern = flash_fwd_kernel_sm100_p4<D, kBf16, kEmu>;
    const DeviceInfo* di = nullptr;
    int rc = current_device_info(&di);
    if (rc != FA_OK) return rc;
    static std::atomic<unsigned long long> attr_mask{0};   // per-device: the attribute belongs to the device's context
    rc = ensure_dynamic_smem(kern, L::kBytes, attr_mask, di->ordinal);
    if (rc != FA_OK) return rc;
    const TileSched ts = make_tile_sched(p);
    const int grid = ts.total < di->num_sms ? ts.total : di->num_sms;
    kern<<<grid, kThreadsP4, L::kBytes, stream>>>(tq, tk, tv, to, kp, ts);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return FA_OK;
}

#define FA_INST(D, B, E)                                                                                              \
    template int launch_fwd_p4<D, B, E>(const fa_fwd_params*, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, \
                                        FwdParams, cudaStream_t);
FA_INST(128, true, 0) FA_INST(128, true, 1) FA_INST(128, true, 2) FA_INST(128, true, 3) FA_INST(128, true, 4)
FA_INST(128, false, 0) FA_INST(128, false, 1) FA_INST(128, false, 2) FA_INST(128, false, 3) FA_INST(128, false, 4)
FA_INST(64, true, 0) FA_INST(64, true, 1) FA_INST(64, true, 2) FA_INST(64, true, 3) FA_INST(64, true, 4)
FA_INST(64, false, 0) FA_INST(64, false, 1) FA_INST(64, false, 2) FA_INST(64, false, 3) FA_INST(64, false, 4)
#undef FA_INST

}  // namespace fa100
