// fa_common.h — host-side helpers shared by the launchers behind the C ABI (include/fa_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../../../include/fa_b200.h"

namespace fa100 {

// thread-local error text + launch counter exposed through fa_b200_last_error / _last_launch_count
void set_error(const char* fmt, ...);
void clear_error();
void count_launch(int n = 1);
void reset_launch_count();

#define FA_CUDA_CHECK(expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            fa100::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FA_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

// Encode a 4-D TMA descriptor over a contiguous [dim3][dim2][dim1][dim0] tensor of 16-bit elements
// (dim0 fastest) with a {box0, box1, box2, box3} tile and 128-byte swizzle.  strides are in BYTES for
// dims 1..3.  Returns FA_OK or FA_ERR_CUDA.
int encode_tmap_4d(CUtensorMap* out, const void* base, bool bf16, const uint64_t dims[4], const uint64_t strides_bytes[3],
                   const uint32_t box[4]);

// Per-device facts, cached per device ordinal (one process may drive several GPUs: cudaFuncSetAttribute, the SM count and
// the compute capability all belong to the CURRENT device, not to the process).
struct DeviceInfo {
    int ordinal;
    int num_sms;
    int cc_major;
};
// fills *out with the entry of the calling thread's current device; FA_OK or FA_ERR_NO_DEVICE / FA_ERR_CUDA
int current_device_info(const DeviceInfo** out);

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: set it the first time `kern` is launched on each device
// (bit `ordinal` of `mask`; devices >= 64 set it on every launch, like the reference's launch templates do)
template <typename Kern>
inline int ensure_dynamic_smem(Kern kern, int bytes, std::atomic<unsigned long long>& mask, int ordinal) {
    if (ordinal >= 0 && ordinal < 64 && ((mask.load(std::memory_order_acquire) >> ordinal) & 1ull)) return FA_OK;
    FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (ordinal >= 0 && ordinal < 64) mask.fetch_or(1ull << ordinal, std::memory_order_release);
    return FA_OK;
}

#ifdef FA_TRACE
// clock64 builds only (make trace): zeroed device buffer [8 roles][64 steps][8 events] the kernels of the next launch
// stamp; read back through fa_b200_trace_read (fa_api.cu)
long long* fa_trace_buffer(cudaStream_t stream);
#endif

// launchers (one translation unit each)
int launch_fwd_sm100(const fa_fwd_params* p, cudaStream_t stream);
int launch_bwd_sm100(const fa_bwd_params* p, cudaStream_t stream);

}  // namespace fa100
