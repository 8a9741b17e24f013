/* fa_b200.h — C ABI of the B200-native fused-attention library (libfa_b200.so).
 *
 * This is the drop-in boundary underneath the reference's operator surface.  The reference
 * (ssiu/flash-attention-turing) packs raw device pointers and sizes into POD structs
 *   Qkv_params / Flash_fwd_params / Flash_bwd_params   (csrc/flash_attn/src/flash.h:6-76)
 * and hands them to
 *   run_mha_fwd_<Headdim,Is_causal>(Flash_fwd_params&)  (flash.h:79)
 *   run_mha_bwd_<Headdim,Is_causal>(Flash_bwd_params&)  (flash.h:82)
 * from mha_fwd / mha_bwd / mha_varlen_fwd / mha_varlen_bwd (csrc/flash_attn/flash_api.cpp:156-468).
 * The two entry points below replace those two launchers; the struct fields mirror flash.h with
 * 64-bit sizes (the reference's int32 offsets, block_info.h:15-21, overflow at >= 2^31 elements)
 * plus an element-type tag (the reference is fp16 only; BASELINE configs are bf16).
 *
 * No torch types cross this boundary: plain pointers, sizes and a cudaStream_t passed as void*.
 * All pointers are DEVICE pointers.  Layouts (contiguous, last index fastest):
 *   fixed length : q,o,dout,dq [b, seqlen_q, h, d]; k,v,dk,dv [b, seqlen_k, h_k, d]
 *   varlen       : q,o,dout,dq [total_q, h, d];     k,v,dk,dv [total_k, h_k, d]
 *                  cu_seqlens_{q,k}: int32[b+1] cumulative row offsets; seqlen_{q,k} = max lengths
 *   lse, dsum    : fp32 [b, h, seqlen_q]   (natural-log LSE of the 1/sqrt(d)-scaled scores;
 *                  0 for rows with no visible key — flash_fwd_kernel.h:766-785)
 * Semantics pinned by the reference: scale 1/sqrt(d); causal mask is bottom-right aligned
 * (keep (i,j) iff j - i <= seqlen_k - seqlen_q, mask.h:20-72); GQA maps q-head hq to kv-head
 * hq / (h / h_k) (flash_fwd_kernel.h:74,91); rows with no visible key produce O = 0, lse = 0.
 */
#ifndef FA_B200_H_
#define FA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_B200_ABI_VERSION 1

typedef enum { FA_DTYPE_FP16 = 0, FA_DTYPE_BF16 = 1 } fa_dtype_t;

/* return codes */
#define FA_OK 0
#define FA_ERR_INVALID_ARG 1    /* unsupported head_dim / dtype / null pointer / bad sizes */
#define FA_ERR_CUDA 2           /* a CUDA runtime/driver call failed; see fa_b200_last_error() */
#define FA_ERR_NO_DEVICE 3      /* not an sm_100 device */

/* mirrors Qkv_params + Flash_fwd_params (flash.h:6-52) */
typedef struct fa_fwd_params {
    const void* q;              /* q_ptr  */
    const void* k;              /* k_ptr  */
    const void* v;              /* v_ptr  */
    void* o;                    /* o_ptr  */
    float* lse;                 /* l_ptr  [b, h, seqlen_q] */
    const int32_t* cu_seqlens_q; /* NULL => fixed length */
    const int32_t* cu_seqlens_k;
    int64_t b;                  /* batch */
    int64_t seqlen_q;           /* (max) query length */
    int64_t seqlen_k;           /* (max) key length */
    int64_t h;                  /* query heads */
    int64_t h_k;                /* key/value heads; h % h_k == 0 */
    int64_t d;                  /* head dim: 64 or 128 */
    int64_t total_q;            /* varlen only: rows in packed q */
    int64_t total_k;            /* varlen only: rows in packed k/v */
    int32_t dtype;              /* fa_dtype_t */
    int32_t is_causal;
} fa_fwd_params;

/* mirrors Flash_bwd_params (flash.h:55-76) */
typedef struct fa_bwd_params {
    fa_fwd_params fwd;          /* q,k,v,o(in),lse(in) + sizes */
    const void* dout;           /* do_ptr */
    void* dq;                   /* dq_ptr */
    void* dk;                   /* dk_ptr  [.., h_k, d]  (GQA group-sum is done inside; the reference */
    void* dv;                   /* dv_ptr   needs h-expanded buffers + torch::sum_out, flash_api.cpp:265-312) */
    float* dsum;                /* do_o_ptr: scratch fp32 [b, h, seqlen_q] = rowsum(dO * O) */
    void* workspace;            /* device scratch of fa_b200_bwd_workspace_bytes() bytes: the fp32 dQ accumulator of the
                                   fused backward (head_dim 128 and 64).  NULL is always allowed: the library then runs the two
                                   deterministic kernels (dQ, dK/dV) that mirror the reference's structure */
} fa_bwd_params;

/* replaces run_mha_fwd (flash_api.cpp:139-145).  Asynchronous on `stream` (a cudaStream_t). */
int fa_b200_fwd(const fa_fwd_params* p, void* stream);

/* replaces run_mha_bwd (flash_api.cpp:147-153): dsum preprocess + dQ + dK/dV on `stream`. */
int fa_b200_bwd(const fa_bwd_params* p, void* stream);

/* bytes of device scratch fa_b200_bwd can use for these sizes (0 if none) */
int64_t fa_b200_bwd_workspace_bytes(const fa_fwd_params* p);

/* number of kernel launches the last fa_b200_fwd / fa_b200_bwd call on this thread enqueued */
int fa_b200_last_launch_count(void);

/* human-readable description of the last error on this thread ("" if none) */
const char* fa_b200_last_error(void);

int fa_b200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FA_B200_H_ */
