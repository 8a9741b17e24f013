// flash_api.cpp — PyTorch operator boundary of the B200 build.
//
// Exposes exactly the reference's pybind surface
//     fwd, bwd, varlen_fwd, varlen_bwd          (/root/reference/csrc/flash_attn/flash_api.cpp:471-476)
// with the same argument order, shapes, return lists and TORCH_CHECK messages (ibid. :156-468), and forwards
// to the C ABI in include/fa_b200.h.  Differences, all deliberate:
//   * fp16 AND bf16 (the reference reinterprets every input as half, :28-31); outputs take q's dtype;
//   * device guard + current stream (the reference launches on the legacy default stream);
//   * dtype / device / contiguity are checked instead of assumed;
//   * outputs are torch::empty where the kernels write every element (the reference memsets 128-512 MiB per
//     call with torch::zeros, :185-188); padding regions that the reference leaves zero are still zero;
//   * GQA dK/dV need no h-expanded scratch nor torch::sum_out (:265-272, :301-312): the group sum is in-kernel.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "../../../include/fa_b200.h"

namespace {

int dtype_tag(const at::Tensor& t) {
    if (t.scalar_type() == at::kHalf) return FA_DTYPE_FP16;
    if (t.scalar_type() == at::kBFloat16) return FA_DTYPE_BF16;
    TORCH_CHECK(false, "q, k, v must be float16 or bfloat16");
    return -1;
}

void check_qkv_common(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v) {
    TORCH_CHECK(q.is_cuda() && k.is_cuda() && v.is_cuda(), "q, k, v must be CUDA tensors");
    TORCH_CHECK(q.scalar_type() == k.scalar_type() && q.scalar_type() == v.scalar_type(), "q, k, v must have the same dtype");
    TORCH_CHECK(q.device() == k.device() && q.device() == v.device(), "q, k, v must be on the same device");
    TORCH_CHECK(q.is_contiguous() && k.is_contiguous() && v.is_contiguous(), "q, k, v must be contiguous");
}

void check_rc(int rc, const char* what) {
    TORCH_CHECK(rc == FA_OK, what, " failed (code ", rc, "): ", fa_b200_last_error());
}

}  // namespace

std::vector<at::Tensor> mha_fwd(at::Tensor q, at::Tensor k, at::Tensor v, bool is_causal) {
    TORCH_CHECK(q.dim() == 4 && k.dim() == 4 && v.dim() == 4, "q, k, v must be rank-4 tensors");
    const int64_t batch_size = q.size(0), seqlen_q = q.size(1), num_heads = q.size(2), head_size = q.size(3);
    const int64_t seqlen_k = k.size(1), num_heads_k = k.size(2);
    TORCH_CHECK(k.size(0) == batch_size && v.size(0) == batch_size, "k/v batch size must match q");
    TORCH_CHECK(v.size(1) == seqlen_k, "k and v seqlen_k must match");
    TORCH_CHECK(v.size(2) == num_heads_k, "k and v num_heads must match");
    TORCH_CHECK(k.size(3) == head_size && v.size(3) == head_size, "q/k/v head_dim must match");
    TORCH_CHECK(num_heads_k > 0 && num_heads % num_heads_k == 0, "num_heads_q must be divisible by num_heads_k for GQA/MQA");
    TORCH_CHECK(head_size == 64 || head_size == 128, "head_dim must be 64 or 128");
    check_qkv_common(q, k, v);
    c10::cuda::CUDAGuard guard(q.device());

    at::Tensor o = torch::empty_like(q);
    at::Tensor l = torch::empty({batch_size, num_heads, seqlen_q}, q.options().dtype(torch::kFloat32));

    fa_fwd_params p = {};
    p.q = q.data_ptr(); p.k = k.data_ptr(); p.v = v.data_ptr(); p.o = o.data_ptr();
    p.lse = l.data_ptr<float>();
    p.b = batch_size; p.seqlen_q = seqlen_q; p.seqlen_k = seqlen_k; p.h = num_heads; p.h_k = num_heads_k; p.d = head_size;
    p.dtype = dtype_tag(q); p.is_causal = is_causal ? 1 : 0;
    check_rc(fa_b200_fwd(&p, at::cuda::getCurrentCUDAStream().stream()), "fa_b200_fwd");
    return {o, l};
}

std::vector<at::Tensor> mha_bwd(at::Tensor q, at::Tensor k, at::Tensor v, at::Tensor out, at::Tensor l, at::Tensor dout,
                                bool is_causal) {
    TORCH_CHECK(q.dim() == 4 && k.dim() == 4 && v.dim() == 4, "q, k, v must be rank-4 tensors");
    TORCH_CHECK(out.dim() == 4 && dout.dim() == 4, "out and dout must be rank-4 tensors");
    const int64_t batch_size = q.size(0), seqlen_q = q.size(1), num_heads = q.size(2), head_size = q.size(3);
    const int64_t seqlen_k = k.size(1), num_heads_k = k.size(2);
    TORCH_CHECK(k.size(0) == batch_size && v.size(0) == batch_size, "k/v batch size must match q");
    TORCH_CHECK(v.size(1) == seqlen_k, "k and v seqlen_k must match");
    TORCH_CHECK(v.size(2) == num_heads_k, "k and v num_heads must match");
    TORCH_CHECK(k.size(3) == head_size && v.size(3) == head_size, "q/k/v head_dim must match");
    TORCH_CHECK(out.sizes() == q.sizes() && dout.sizes() == q.sizes(), "out and dout must match q shape");
    TORCH_CHECK(num_heads_k > 0 && num_heads % num_heads_k == 0, "num_heads_q must be divisible by num_heads_k for GQA/MQA");
    TORCH_CHECK(head_size == 64 || head_size == 128, "head_dim must be 64 or 128");
    check_qkv_common(q, k, v);
    TORCH_CHECK(out.is_cuda() && dout.is_cuda() && l.is_cuda(), "out, dout, l must be CUDA tensors");
    TORCH_CHECK(out.scalar_type() == q.scalar_type() && dout.scalar_type() == q.scalar_type(), "out and dout must have q's dtype");
    TORCH_CHECK(l.scalar_type() == at::kFloat && l.dim() == 3 && l.size(0) == batch_size && l.size(1) == num_heads &&
                    l.size(2) == seqlen_q, "l must be float32 with shape [batch_size, nheads_q, seqlen_q]");
    c10::cuda::CUDAGuard guard(q.device());
    out = out.contiguous(); dout = dout.contiguous(); l = l.contiguous();

    at::Tensor dq = torch::empty_like(q);
    at::Tensor dk = torch::empty_like(k);
    at::Tensor dv = torch::empty_like(v);
    at::Tensor do_o = torch::empty_like(l);

    fa_bwd_params p = {};
    p.fwd.q = q.data_ptr(); p.fwd.k = k.data_ptr(); p.fwd.v = v.data_ptr(); p.fwd.o = out.data_ptr();
    p.fwd.lse = l.data_ptr<float>();
    p.fwd.b = batch_size; p.fwd.seqlen_q = seqlen_q; p.fwd.seqlen_k = seqlen_k; p.fwd.h = num_heads; p.fwd.h_k = num_heads_k;
    p.fwd.d = head_size; p.fwd.dtype = dtype_tag(q); p.fwd.is_causal = is_causal ? 1 : 0;
    p.dout = dout.data_ptr(); p.dq = dq.data_ptr(); p.dk = dk.data_ptr(); p.dv = dv.data_ptr(); p.dsum = do_o.data_ptr<float>();
    at::Tensor ws;
    const int64_t ws_bytes = fa_b200_bwd_workspace_bytes(&p.fwd);
    if (ws_bytes > 0) { ws = torch::empty({ws_bytes}, q.options().dtype(torch::kUInt8)); p.workspace = ws.data_ptr(); }
    check_rc(fa_b200_bwd(&p, at::cuda::getCurrentCUDAStream().stream()), "fa_b200_bwd");
    return {dq, dk, dv};
}

static void check_varlen_common(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& cu_seqlens_q,
                                const at::Tensor& cu_seqlens_k) {
    TORCH_CHECK(q.dim() == 3 && k.dim() == 3 && v.dim() == 3, "q, k, v must be rank-3 packed tensors");
    TORCH_CHECK(cu_seqlens_q.is_cuda() && cu_seqlens_k.is_cuda(), "cu_seqlens_q/cu_seqlens_k must be CUDA tensors");
    TORCH_CHECK(cu_seqlens_q.scalar_type() == torch::kInt32 && cu_seqlens_k.scalar_type() == torch::kInt32,
                "cu_seqlens_q/cu_seqlens_k must be int32 tensors");
    TORCH_CHECK(cu_seqlens_q.is_contiguous() && cu_seqlens_k.is_contiguous(), "cu_seqlens_q/cu_seqlens_k must be contiguous");
    TORCH_CHECK(cu_seqlens_q.dim() == 1 && cu_seqlens_k.dim() == 1, "cu_seqlens_q/cu_seqlens_k must be rank-1");
    TORCH_CHECK(cu_seqlens_q.numel() >= 2 && cu_seqlens_k.numel() >= 2, "cu_seqlens_q/cu_seqlens_k must have at least 2 elements");
    TORCH_CHECK(cu_seqlens_k.numel() == cu_seqlens_q.numel(), "cu_seqlens_k must have shape [batch_size + 1] with cumulative offsets");
    TORCH_CHECK(k.size(0) == v.size(0), "k and v total tokens must match");
    TORCH_CHECK(k.size(1) == v.size(1), "k and v num_heads must match");
    TORCH_CHECK(k.size(2) == v.size(2), "k and v head_dim must match");
    TORCH_CHECK(q.size(2) == k.size(2), "q/k/v head_dim must match");
    TORCH_CHECK(k.size(1) > 0 && q.size(1) % k.size(1) == 0, "num_heads_q must be divisible by num_heads_k for GQA/MQA");
    TORCH_CHECK(q.size(2) == 64 || q.size(2) == 128, "head_dim must be 64 or 128");
    check_qkv_common(q, k, v);
}

std::vector<at::Tensor> mha_varlen_fwd(at::Tensor q, at::Tensor k, at::Tensor v, at::Tensor& cu_seqlens_q,
                                       at::Tensor& cu_seqlens_k, const int max_seqlen_q, const int max_seqlen_k,
                                       bool is_causal) {
    check_varlen_common(q, k, v, cu_seqlens_q, cu_seqlens_k);
    const int64_t batch_size = cu_seqlens_q.numel() - 1;
    const int64_t num_heads = q.size(1), num_heads_k = k.size(1), head_size = q.size(2);
    c10::cuda::CUDAGuard guard(q.device());

    // rows outside [cu_seqlens[0], cu_seqlens[b]) and the l padding beyond each sequence stay zero, as in the
    // reference (torch::zeros_like / torch::zeros, flash_api.cpp:351-352)
    at::Tensor out = torch::zeros_like(q);
    at::Tensor l = torch::zeros({batch_size, num_heads, (int64_t)max_seqlen_q}, q.options().dtype(torch::kFloat32));

    fa_fwd_params p = {};
    p.q = q.data_ptr(); p.k = k.data_ptr(); p.v = v.data_ptr(); p.o = out.data_ptr(); p.lse = l.data_ptr<float>();
    p.cu_seqlens_q = cu_seqlens_q.data_ptr<int32_t>(); p.cu_seqlens_k = cu_seqlens_k.data_ptr<int32_t>();
    p.b = batch_size; p.seqlen_q = max_seqlen_q; p.seqlen_k = max_seqlen_k; p.h = num_heads; p.h_k = num_heads_k; p.d = head_size;
    p.total_q = q.size(0); p.total_k = k.size(0);
    p.dtype = dtype_tag(q); p.is_causal = is_causal ? 1 : 0;
    check_rc(fa_b200_fwd(&p, at::cuda::getCurrentCUDAStream().stream()), "fa_b200_fwd");
    return {out, l};
}

std::vector<at::Tensor> mha_varlen_bwd(at::Tensor q, at::Tensor k, at::Tensor v, at::Tensor out, at::Tensor l, at::Tensor dout,
                                       at::Tensor cu_seqlens_q, at::Tensor cu_seqlens_k, const int max_seqlen_q,
                                       const int max_seqlen_k, bool is_causal) {
    check_varlen_common(q, k, v, cu_seqlens_q, cu_seqlens_k);
    const int64_t batch_size = cu_seqlens_q.numel() - 1;
    TORCH_CHECK(out.sizes() == q.sizes(), "out must match q shape");
    TORCH_CHECK(dout.sizes() == q.sizes(), "dout must match q shape");
    TORCH_CHECK(l.dim() == 3, "l must be rank-3 for varlen_bwd");
    const int64_t num_heads = q.size(1), num_heads_k = k.size(1), head_size = q.size(2);
    TORCH_CHECK(l.size(0) == batch_size && l.size(1) == num_heads && l.size(2) == max_seqlen_q,
                "l must have shape [batch_size, nheads_q, max_seqlen_q]");
    TORCH_CHECK(out.is_cuda() && dout.is_cuda() && l.is_cuda(), "out, dout, l must be CUDA tensors");
    TORCH_CHECK(out.scalar_type() == q.scalar_type() && dout.scalar_type() == q.scalar_type(), "out and dout must have q's dtype");
    TORCH_CHECK(l.scalar_type() == at::kFloat, "l must be float32");
    c10::cuda::CUDAGuard guard(q.device());
    out = out.contiguous(); dout = dout.contiguous(); l = l.contiguous();

    at::Tensor dq = torch::zeros_like(q);
    at::Tensor dk = torch::zeros_like(k);
    at::Tensor dv = torch::zeros_like(v);
    at::Tensor do_o = torch::zeros_like(l);

    fa_bwd_params p = {};
    p.fwd.q = q.data_ptr(); p.fwd.k = k.data_ptr(); p.fwd.v = v.data_ptr(); p.fwd.o = out.data_ptr();
    p.fwd.lse = l.data_ptr<float>();
    p.fwd.cu_seqlens_q = cu_seqlens_q.data_ptr<int32_t>(); p.fwd.cu_seqlens_k = cu_seqlens_k.data_ptr<int32_t>();
    p.fwd.b = batch_size; p.fwd.seqlen_q = max_seqlen_q; p.fwd.seqlen_k = max_seqlen_k; p.fwd.h = num_heads; p.fwd.h_k = num_heads_k;
    p.fwd.d = head_size; p.fwd.total_q = q.size(0); p.fwd.total_k = k.size(0);
    p.fwd.dtype = dtype_tag(q); p.fwd.is_causal = is_causal ? 1 : 0;
    p.dout = dout.data_ptr(); p.dq = dq.data_ptr(); p.dk = dk.data_ptr(); p.dv = dv.data_ptr(); p.dsum = do_o.data_ptr<float>();
    at::Tensor ws;
    const int64_t ws_bytes = fa_b200_bwd_workspace_bytes(&p.fwd);
    if (ws_bytes > 0) { ws = torch::empty({ws_bytes}, q.options().dtype(torch::kUInt8)); p.workspace = ws.data_ptr(); }
    check_rc(fa_b200_bwd(&p, at::cuda::getCurrentCUDAStream().stream()), "fa_b200_bwd");
    return {dq, dk, dv};
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("fwd", &mha_fwd, "Forward pass");
    m.def("bwd", &mha_bwd, "Backward pass");
    m.def("varlen_fwd", &mha_varlen_fwd, "Varlen forward pass");
    m.def("varlen_bwd", &mha_varlen_bwd, "Varlen backward pass");
    m.def("last_launch_count", []() { return fa_b200_last_launch_count(); }, "kernel launches of the last fwd/bwd call");
    m.def("abi_version", []() { return fa_b200_abi_version(); });
}
