// sm100_ptx.cuh — thin inline-PTX layer for Blackwell (sm_100a): mbarrier, TMA, tcgen05 / TMEM.
//
// Everything the attention kernels need from the hardware is spelled out here as raw PTX; there is
// no CUTLASS/CuTe dependency anywhere in this tree.  (The reference gets its tensor-core and copy
// atoms from CuTe's sm_75 layer — /root/reference/csrc/flash_attn/src/kernel_traits.h:22-38 — which
// this file replaces wholesale.)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fa100 {

#define FA_DEVICE __device__ __forceinline__

// ------------------------------------------------------------------------------------------------
// address helpers
// ------------------------------------------------------------------------------------------------
FA_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// Every helper below that names a shared-memory location takes either a generic pointer or a 32-bit address in the shared
// window.  Kernels that keep only 32-bit addresses never hold the 64-bit generic base pointer: in the forward kernel it was
// the one value ptxas spilled, and with all of L1 configured as shared memory every reload was an L2 round trip.
FA_DEVICE uint32_t smem_u32(uint32_t shared_window_address) { return shared_window_address; }

FA_DEVICE bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
template <class B> FA_DEVICE void mbar_init(B bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
FA_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
FA_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <class B> FA_DEVICE void mbar_arrive(B bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <class B> FA_DEVICE void mbar_arrive_expect_tx(B bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
template <class B> FA_DEVICE bool mbar_try_wait(B bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking probe of a phase (test_wait never suspends the thread).
template <class B> FA_DEVICE bool mbar_test_wait(B bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Blocking wait on phase parity.  try_wait is a hardware-suspended wait with a time limit, so the loop
// spins at most a few times.  With FA_HANG_GUARD the wait traps after ~2^28 polls instead of hanging the
// GPU box (used in bring-up builds).
#ifdef FA_HANG_GUARD
// diagnosis builds: the first wait that exceeds ~0.1 s records who waited on what and every wait of the process gives up
// (results are garbage, the kernel ends); read back with fa_b200_hang_read() (flash_bwd_tc_sm100.cu)
static __device__ unsigned int g_hang_flag;
static __device__ unsigned int g_hang_info[8];
#endif
template <class B> FA_DEVICE void mbar_wait(B bar, uint32_t parity) {
#ifdef FA_HANG_GUARD
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (*(volatile unsigned int*)&g_hang_flag) return;
        if (clock64() - t0 > 200000000LL) {
            if (atomicExch(&g_hang_flag, 1u) == 0u) {
                g_hang_info[0] = smem_u32(bar); g_hang_info[1] = parity; g_hang_info[2] = threadIdx.x;
                g_hang_info[3] = blockIdx.x; g_hang_info[4] = blockIdx.y; g_hang_info[5] = blockIdx.z;
            }
            return;
        }
    }
#else
    while (!mbar_try_wait(bar, parity)) {}
#endif
}

// ------------------------------------------------------------------------------------------------
// named barriers (sub-CTA sync between the warps of one role)
// ------------------------------------------------------------------------------------------------
FA_DEVICE void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
FA_DEVICE void named_bar_arrive(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int N> FA_DEVICE void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> FA_DEVICE void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 4-D tiled loads / stores.  Coordinates are innermost-first.
// ------------------------------------------------------------------------------------------------
FA_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
template <class D, class B> FA_DEVICE void tma_load_4d(D smem_dst, const CUtensorMap* m, B bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
        : "memory");
}
template <class D, class B> FA_DEVICE void tma_load_4d_hint(D smem_dst, const CUtensorMap* m, B bar, int c0, int c1, int c2,
                                int c3, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3), "l"(policy)
        : "memory");
}
template <class S> FA_DEVICE void tma_store_4d(const CUtensorMap* m, S smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
FA_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> FA_DEVICE void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> FA_DEVICE void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// L2 cache-policy words (same encodings the driver's createpolicy produces for the fractional 1.0 case)
constexpr uint64_t kPolicyEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kPolicyEvictLast = 0x14F0000000000000ull;
constexpr uint64_t kPolicyEvictNormal = 0x1000000000000000ull;

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ------------------------------------------------------------------------------------------------
template <int NCOLS, class R> FA_DEVICE void tmem_alloc(R smem_result) {  // whole warp, .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(NCOLS)
                 : "memory");
}
FA_DEVICE void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
template <int NCOLS> FA_DEVICE void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

FA_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
FA_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
FA_DEVICE void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
FA_DEVICE void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.commit: arrive (count 1) on an mbarrier once all previously issued tcgen05.mma of this thread
// have completed.  Implies tcgen05.fence::before_thread_sync.
template <class B> FA_DEVICE void tc_commit(B bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: descriptors
// ------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4          [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4 [46,48) descriptor version (1 on sm_100)
//   [49,52) base offset (0: tiles are 1024-B aligned)   [61,64) swizzle mode (2 = 128 B)
// K-major, 128B swizzle : rows are 128 B (64 x 16-bit) long, 8-row groups SBO apart; LBO unused (=1).
// MN-major, 128B swizzle: 64 MN-contiguous elements per 128 B row; the 8 K-rows of a group are 128 B
//                         apart, K-groups are SBO apart, successive 64-element MN blocks are LBO apart.
constexpr uint64_t kDescSwizzle128 = (uint64_t(2) << 61) | (uint64_t(1) << 46);
FA_DEVICE uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return kDescSwizzle128 | uint64_t((smem_addr & 0x3FFFF) >> 4) | (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16) |
           (uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32);
}

// Split form used on the MMA issue path: the low word carries (address >> 4) and the leading-dim offset, the high
// word (stride offset 1024 B, version, 128B swizzle) is a compile-time constant; stepping a descriptor is then a
// single 32-bit add on the low word (tile bases are 1024-B aligned and shared memory is < 256 KB, so no carry).
constexpr uint32_t kDescHiK = uint32_t((kDescSwizzle128 >> 32) | (1024u >> 4));
FA_DEVICE uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
FA_DEVICE uint64_t desc_make(uint32_t lo, uint32_t hi) { return (uint64_t(hi) << 32) | lo; }

// Instruction descriptor for kind::f16 (fp16/bf16 operands, fp32 accumulate):
//   [4,6) D format (1 = f32)  [7,10) A format  [10,13) B format (0 = f16, 1 = bf16)
//   [15] A major  [16] B major (0 = K-major, 1 = MN-major)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(bool bf16, int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((a_mn_major ? 1u : 0u) << 15) |
           ((b_mn_major ? 1u : 0u) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
FA_DEVICE void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: 128 lanes x K/2 32-bit columns, two 16-bit values per column)
FA_DEVICE void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM <-> registers.  Shape 32x32b: thread i of the warp owns TMEM lane (lane_base + i) and
// receives N consecutive 32-bit columns.  A warp may only touch lanes [32*(warp_id%4), +32).
// ------------------------------------------------------------------------------------------------
FA_DEVICE void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
FA_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
FA_DEVICE void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
FA_DEVICE void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
FA_DEVICE void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// ------------------------------------------------------------------------------------------------
// explicit shared-space accesses (the dynamic smem pointer goes through integer alignment arithmetic, after which the
// compiler only knows a generic address and would emit LD.E / ST.E through the generic path)
// ------------------------------------------------------------------------------------------------
FA_DEVICE float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
FA_DEVICE uint4 lds128u(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
FA_DEVICE void sts128u(uint32_t saddr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
FA_DEVICE int lds32(uint32_t saddr) { int v; asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory"); return v; }
FA_DEVICE void sts32(uint32_t saddr, int v) { asm volatile("st.volatile.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
FA_DEVICE int atoms_add(uint32_t saddr, int v) { int o; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(saddr), "r"(v) : "memory"); return o; }
FA_DEVICE int atoms_cas(uint32_t saddr, int cmp, int v) { int o; asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(o) : "r"(saddr), "r"(cmp), "r"(v) : "memory"); return o; }
FA_DEVICE int atoms_exch(uint32_t saddr, int v) { int o; asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(o) : "r"(saddr), "r"(v) : "memory"); return o; }
FA_DEVICE void sts32f(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }

// ------------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------------
FA_DEVICE float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x for a pair of floats on the FMA pipe (no MUFU): Cody-Waite split x = n + f, f in [-0.5, 0.5], degree-3 minimax
// polynomial for 2^f (|rel err| < 7.5e-5, far below 16-bit rounding), exponent inserted with an integer add.
// x is clamped to [-125, 126] (NaN / -inf -> 2^-125, i.e. zero after any 16-bit rounding; above 126 the exponent add
// would wrap around).
FA_DEVICE float2 exp2_poly_pair(float2 x) {
    const float2 magic = make_float2(12582912.f, 12582912.f);                 // 1.5 * 2^23: low mantissa bits of x + magic = rint(x)
    x = make_float2(fminf(fmaxf(x.x, -125.f), 126.f), fminf(fmaxf(x.y, -125.f), 126.f));
    const float2 tt = __fadd2_rn(x, magic);
    const float2 nnf = __ffma2_rn(tt, make_float2(-1.f, -1.f), magic);        // -rint(x), exact
    const float2 f = __fadd2_rn(x, nnf);
    float2 pl = __ffma2_rn(make_float2(0.05517115816473961f, 0.05517115816473961f), f,
                           make_float2(0.2426101416349411f, 0.2426101416349411f));
    pl = __ffma2_rn(pl, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
    pl = __ffma2_rn(pl, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
    return make_float2(__uint_as_float(__float_as_uint(pl.x) + (__float_as_uint(tt.x) << 23)),
                       __uint_as_float(__float_as_uint(pl.y) + (__float_as_uint(tt.y) << 23)));
}
// pack two fp32 into one 32-bit word of two 16-bit floats: lo -> bits [0,16), hi -> bits [16,32)
template <bool kBf16> FA_DEVICE uint32_t pack2(float lo, float hi) {
    uint32_t r;
    if constexpr (kBf16) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    } else {
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    }
    return r;
}

}  // namespace fa100
