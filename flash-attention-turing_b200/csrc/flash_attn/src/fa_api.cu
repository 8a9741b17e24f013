// fa_api.cu — implementation of the C ABI declared in include/fa_b200.h.
//
// Replaces the reference's runtime->template dispatch (run_mha_fwd / run_mha_bwd,
// /root/reference/csrc/flash_attn/flash_api.cpp:139-153 + static_switch.h:18-38).  Unlike the
// reference's HEADDIM_SWITCH, an unsupported head_dim is an error here, not a silent no-op.
#include <stdarg.h>
#include <string.h>

#include "fa_common.h"
#include "flash_bwd_params.h"

namespace fa100 {

static thread_local char g_err[512] = {0};
static thread_local int g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void clear_error() { g_err[0] = 0; }
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved through the runtime so that libfa_b200.so has no link-time dependency on libcuda.so
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
        set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
        return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(sym);
    return fn;
}

int encode_tmap_4d(CUtensorMap* out, const void* base, bool bf16, const uint64_t dims[4], const uint64_t strides_bytes[3],
                   const uint32_t box[4]) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return FA_ERR_CUDA;
    cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gstr[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                    const_cast<void*>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %llu,%llu,%llu,%llu base %p)", (int)r,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                  (unsigned long long)dims[3], base);
        return FA_ERR_CUDA;
    }
    return FA_OK;
}

int current_device_info(const DeviceInfo** out) {
    constexpr int kMaxDev = 64;
    static DeviceInfo table[kMaxDev];
    static std::atomic<int> ready[kMaxDev];          // 0 unknown, 1 filled (zero-initialised)
    static thread_local DeviceInfo overflow;         // ordinals >= kMaxDev: queried on every call
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        set_error("no CUDA device");
        return FA_ERR_NO_DEVICE;
    }
    DeviceInfo* di = (dev >= 0 && dev < kMaxDev) ? &table[dev] : &overflow;
    if (dev < 0 || dev >= kMaxDev || ready[dev].load(std::memory_order_acquire) == 0) {
        DeviceInfo tmp;
        tmp.ordinal = dev;
        if (cudaDeviceGetAttribute(&tmp.cc_major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&tmp.num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            set_error("cannot query device %d", dev);
            return FA_ERR_CUDA;
        }
        *di = tmp;                                    // racing threads write identical values
        if (dev >= 0 && dev < kMaxDev) ready[dev].store(1, std::memory_order_release);
    }
    if (di->cc_major != 10) {
        set_error("libfa_b200 needs an sm_100 (Blackwell B200) device, device %d is sm_%d0; there is no fallback path", dev, di->cc_major);
        return FA_ERR_NO_DEVICE;
    }
    *out = di;
    return FA_OK;
}

static int check_device() {
    const DeviceInfo* di = nullptr;
    return current_device_info(&di);
}

static int validate_fwd(const fa_fwd_params* p) {
    if (!p) { set_error("null params"); return FA_ERR_INVALID_ARG; }
    if (p->b < 0 || p->seqlen_q < 0 || p->seqlen_k < 0 || p->h <= 0 || p->h_k <= 0) {
        set_error("negative / zero sizes"); return FA_ERR_INVALID_ARG;
    }
    if (p->h % p->h_k != 0) { set_error("num_heads_q must be divisible by num_heads_k for GQA/MQA"); return FA_ERR_INVALID_ARG; }
    if (p->d != 64 && p->d != 128) { set_error("head_dim %lld not supported (64 or 128)", (long long)p->d); return FA_ERR_INVALID_ARG; }
    if (p->dtype != FA_DTYPE_FP16 && p->dtype != FA_DTYPE_BF16) { set_error("dtype must be fp16 or bf16"); return FA_ERR_INVALID_ARG; }
    if ((p->cu_seqlens_q == nullptr) != (p->cu_seqlens_k == nullptr)) { set_error("cu_seqlens_q and cu_seqlens_k must both be set or both be NULL"); return FA_ERR_INVALID_ARG; }
    const bool empty = p->b == 0 || p->seqlen_q == 0 || (p->cu_seqlens_q && p->total_q == 0);
    if (!empty && (!p->q || !p->o || !p->lse)) { set_error("null q/o/lse pointer"); return FA_ERR_INVALID_ARG; }
    if (!empty && p->seqlen_k > 0 && (!p->k || !p->v)) { set_error("null k/v pointer"); return FA_ERR_INVALID_ARG; }
    if (p->seqlen_q > INT32_MAX || p->seqlen_k > INT32_MAX || p->b > 65535 || p->h > 65535 ||
        p->total_q > INT32_MAX || p->total_k > INT32_MAX) {
        set_error("size out of range"); return FA_ERR_INVALID_ARG;
    }
    // the device-side work lists index (256-row query block, batch, head) triples and (batch * head) pairs with 32 bits
    if (p->b * p->h > INT32_MAX / 2 || ((p->seqlen_q + 63) / 64) * p->b * p->h > INT32_MAX) {
        set_error("batch * heads * query blocks exceeds the 32-bit work list"); return FA_ERR_INVALID_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(p->q) | reinterpret_cast<uintptr_t>(p->k) | reinterpret_cast<uintptr_t>(p->v) |
         reinterpret_cast<uintptr_t>(p->o)) & 15) {
        set_error("q/k/v/o must be 16-byte aligned"); return FA_ERR_INVALID_ARG;
    }
    return FA_OK;
}

#ifdef FA_TRACE
static long long* g_trace = nullptr;
constexpr size_t kTraceWords = 8 * 64 * 8;
long long* fa_trace_buffer(cudaStream_t stream) {
    if (!g_trace && cudaMalloc(&g_trace, kTraceWords * sizeof(long long)) != cudaSuccess) return nullptr;
    cudaMemsetAsync(g_trace, 0, kTraceWords * sizeof(long long), stream);
    return g_trace;
}
#endif

}  // namespace fa100

using namespace fa100;

#ifdef FA_TRACE
extern "C" int fa_b200_trace_read(long long* host, int words) {   // trace builds only; not part of include/fa_b200.h
    if (!g_trace || words > (int)kTraceWords) return -1;
    cudaDeviceSynchronize();
    return cudaMemcpy(host, g_trace, (size_t)words * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" {

int fa_b200_abi_version(void) { return FA_B200_ABI_VERSION; }
const char* fa_b200_last_error(void) { return g_err; }
int fa_b200_last_launch_count(void) { return g_launches; }

int fa_b200_fwd(const fa_fwd_params* p, void* stream) {
    clear_error();
    reset_launch_count();
    int rc = validate_fwd(p);
    if (rc != FA_OK) return rc;
    rc = check_device();
    if (rc != FA_OK) return rc;
    return launch_fwd_sm100(p, static_cast<cudaStream_t>(stream));
}

int64_t fa_b200_bwd_workspace_bytes(const fa_fwd_params* p) {
    if (!p) return 0;
    return bwd_fused_workspace_bytes(p->b, p->seqlen_q, p->h, p->d);   // 0 unless FA_B200_BWD=fused (fp32 dQ accumulator)
}

int fa_b200_bwd(const fa_bwd_params* p, void* stream) {
    clear_error();
    reset_launch_count();
    if (!p) { set_error("null params"); return FA_ERR_INVALID_ARG; }
    int rc = validate_fwd(&p->fwd);
    if (rc != FA_OK) return rc;
    const fa_fwd_params* f = &p->fwd;
    const bool empty = f->b == 0 || f->seqlen_q == 0 || (f->cu_seqlens_q && f->total_q == 0);
    if (!empty && (!p->dout || !p->dq || !p->dsum)) { set_error("null dout/dq/dsum pointer"); return FA_ERR_INVALID_ARG; }
    if (!empty && f->seqlen_k > 0 && (!p->dk || !p->dv)) { set_error("null dk/dv pointer"); return FA_ERR_INVALID_ARG; }
    rc = check_device();
    if (rc != FA_OK) return rc;
    return launch_bwd_sm100(p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
