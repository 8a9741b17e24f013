// Micro-benchmark (GPU box): tcgen05.mma issue/execute cost per instruction for the shapes the attention kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flash-attention-turing_b200/csrc/flash_attn/src -o scripts/ubench_umma.bin scripts/ubench_umma.cu
#include <cstdio>
#include "sm100_ptx.cuh"
using namespace fa100;

// mode 0: SS K-major A,B   1: TS (A from TMEM), B K-major   2: TS, B MN-major   3: SS, B MN-major
template <int N, int MODE, int CHAINS>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x < 32) { tmem_alloc<512>(&slot); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t a_lo = desc_lo(smem_u32(smem), 16);
        const uint32_t b_lo = desc_lo(smem_u32(smem + 32768), (MODE == 2 || MODE == 3) ? 16384 : 16);
        const uint32_t idesc = make_idesc(true, 128, N, false, MODE == 2 || MODE == 3);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) {
                    const uint32_t off = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
                    const uint32_t boff = (MODE == 2 || MODE == 3) ? (kk * 2048) >> 4 : off;
                    if (MODE == 0 || MODE == 3) umma_ss(tm + c * 256, desc_make(a_lo + off, kDescHiK), desc_make(b_lo + boff, kDescHiK), idesc, kk > 0);
                    else umma_ts(tm + c * 256, tm + 128 + kk * 8, desc_make(b_lo + boff, kDescHiK), idesc, kk > 0);
                }
            }
        }
        const long long t1 = clock64();
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int N, int MODE, int CHAINS> void run(const char* name) {
    long long* d; cudaMalloc(&d, 148 * 2 * 8);
    cudaFuncSetAttribute(bench<N, MODE, CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 64;
    bench<N, MODE, CHAINS><<<148, 128, 100 * 1024>>>(d, iters);
    bench<N, MODE, CHAINS><<<148, 128, 100 * 1024>>>(d, iters);
    cudaDeviceSynchronize();
    long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double a = 0, b = 0; for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
    const double n = 148.0 * iters * 8 * CHAINS;
    printf("UMMA M128 N%-3d K16 %-28s chains %d : issue %6.1f cyc/MMA, issue+drain %6.1f cyc/MMA (ideal %d) %s\n", N, name, CHAINS, a / n, b / n,
           128 * N / 256, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    run<128, 0, 1>("SS  A,B K-major");
    run<128, 0, 2>("SS  A,B K-major");
    run<64, 0, 1>("SS  A,B K-major");
    run<64, 0, 2>("SS  A,B K-major");
    run<32, 0, 2>("SS  A,B K-major");
    run<256, 0, 1>("SS  A,B K-major");
    run<128, 1, 1>("TS  B K-major");
    run<64, 1, 2>("TS  B K-major");
    run<128, 2, 1>("TS  B MN-major");
    run<128, 2, 2>("TS  B MN-major");
    run<64, 2, 2>("TS  B MN-major");
    run<128, 3, 1>("SS  B MN-major");
    return 0;
}
