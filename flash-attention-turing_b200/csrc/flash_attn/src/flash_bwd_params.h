// flash_bwd_params.h — device-side parameter block shared by the backward kernels (mirrors Flash_bwd_params,
// /root/reference/csrc/flash_attn/src/flash.h:55-76, with 64-bit-safe indexing done in the kernels).
#pragma once
#include <cuda_runtime.h>

namespace fa100 {

struct BwdParams {
    const void *q, *k, *v, *o, *dout;
    const float* lse;
    float* dsum;
    void *dq, *dk, *dv;
    const int* cu_q;
    const int* cu_k;
    int b, sq, sk, h, h_k, hratio, d;
    int is_causal;
    float scale;
    long long total_q, total_k;   // varlen: rows of the packed q / k tensors
    float* dqacc;                 // fused backward only: fp32 dQ accumulator [b][h][sq_pad][d] (the C ABI's workspace)
    int sq_pad;                   // sq rounded up to a multiple of 64
    long long* trace;             // FA_TRACE builds only
};

#ifdef FA_TRACE
#define FA_BTRACE(role, j, ev)                                                                   \
    do {                                                                                         \
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 64)        \
            p.trace[((role) * 64 + (j)) * 8 + (ev)] = clock64();                                 \
    } while (0)
#else
#define FA_BTRACE(role, j, ev) do {} while (0)
#endif

// tensor-core (tcgen05) backward: dQ kernel + dK/dV kernel.  Returns FA_OK, or a negative value if this shape is
// not handled (never happens for d in {64,128}).
int launch_bwd_tc_sm100(const BwdParams& kp, bool bf16, cudaStream_t stream);

// Fused backward (default; FA_B200_BWD=det switches it off): one K/V-stationary kernel produces dK, dV and — through fp32
// bulk reductions into the workspace — dQ.  head_dim 128 and 64 (FA_B200_BWD_D64=det keeps 64 on the two kernels); needs a workspace of bwd_fused_workspace_bytes(); a caller
// that passes no workspace gets the deterministic two-kernel path.
bool bwd_fused_enabled();
long long bwd_fused_workspace_bytes(long long b, long long sq_max, long long h, long long d);

}  // namespace fa100
