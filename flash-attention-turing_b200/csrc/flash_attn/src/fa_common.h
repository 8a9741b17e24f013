// fa_common.h — host-side helpers shared by the launchers behind the C ABI (include/fa_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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

// launchers (one translation unit each)
int launch_fwd_sm100(const fa_fwd_params* p, cudaStream_t stream);
int launch_bwd_sm100(const fa_bwd_params* p, cudaStream_t stream);

}  // namespace fa100
