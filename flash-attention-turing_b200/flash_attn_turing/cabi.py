"""ctypes view of include/fa_b200.h: the binding a reference maintainer would write to call libfa_b200.so without going
through torch's pybind layer.  bench.py and the GPU parity tests call the kernels through it.
FA_B200_LIB overrides the library path (A/B runs against another build of the same ABI)."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_HERE))
LIB_PATH = os.environ.get("FA_B200_LIB") or os.path.join(_HERE, "libfa_b200.so")
HEADER = os.path.join(ROOT, "include", "fa_b200.h")

FA_OK, FA_ERR_INVALID_ARG, FA_ERR_CUDA, FA_ERR_NO_DEVICE = 0, 1, 2, 3
FA_DTYPE_FP16, FA_DTYPE_BF16 = 0, 1


class FwdParams(ctypes.Structure):
    _fields_ = [("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p), ("o", ctypes.c_void_p),
                ("lse", ctypes.c_void_p), ("cu_seqlens_q", ctypes.c_void_p), ("cu_seqlens_k", ctypes.c_void_p),
                ("b", ctypes.c_int64), ("seqlen_q", ctypes.c_int64), ("seqlen_k", ctypes.c_int64),
                ("h", ctypes.c_int64), ("h_k", ctypes.c_int64), ("d", ctypes.c_int64),
                ("total_q", ctypes.c_int64), ("total_k", ctypes.c_int64),
                ("dtype", ctypes.c_int32), ("is_causal", ctypes.c_int32)]


class BwdParams(ctypes.Structure):
    _fields_ = [("fwd", FwdParams), ("dout", ctypes.c_void_p), ("dq", ctypes.c_void_p), ("dk", ctypes.c_void_p),
                ("dv", ctypes.c_void_p), ("dsum", ctypes.c_void_p), ("workspace", ctypes.c_void_p)]


def declared_symbols():
    """every function include/fa_b200.h declares"""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fa_b200_\w+)\s*\(", text)))


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(LIB_PATH)
        lib.fa_b200_fwd.argtypes = [ctypes.POINTER(FwdParams), ctypes.c_void_p]
        lib.fa_b200_fwd.restype = ctypes.c_int
        lib.fa_b200_bwd.argtypes = [ctypes.POINTER(BwdParams), ctypes.c_void_p]
        lib.fa_b200_bwd.restype = ctypes.c_int
        lib.fa_b200_bwd_workspace_bytes.argtypes = [ctypes.POINTER(FwdParams)]
        lib.fa_b200_bwd_workspace_bytes.restype = ctypes.c_int64
        lib.fa_b200_last_error.restype = ctypes.c_char_p
        lib.fa_b200_last_launch_count.restype = ctypes.c_int
        lib.fa_b200_abi_version.restype = ctypes.c_int
        _lib = lib
    return _lib


def _dtype_tag(t):
    import torch
    return {torch.float16: FA_DTYPE_FP16, torch.bfloat16: FA_DTYPE_BF16}[t.dtype]


def make_fwd_params(q, k, v, o, lse, causal, cu_q=None, cu_k=None, max_sq=None, max_sk=None):
    p = FwdParams()
    p.q, p.k, p.v, p.o, p.lse = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr()
    if cu_q is None:
        p.b, p.seqlen_q, p.h, p.d = q.shape
        p.seqlen_k, p.h_k = k.shape[1], k.shape[2]
    else:
        p.cu_seqlens_q, p.cu_seqlens_k = cu_q.data_ptr(), cu_k.data_ptr()
        p.b = cu_q.numel() - 1
        p.seqlen_q, p.seqlen_k = int(max_sq), int(max_sk)
        p.total_q, p.h, p.d = q.shape
        p.total_k, p.h_k = k.shape[0], k.shape[1]
    p.dtype = _dtype_tag(q)
    p.is_causal = int(bool(causal))
    return p


def fwd(q, k, v, causal, cu_q=None, cu_k=None, max_sq=None, max_sk=None):
    """torch CUDA tensors in, (o, lse) out — through fa_b200_fwd on the current stream"""
    import torch
    lib = load()
    if cu_q is None:
        o = torch.empty_like(q)
        lse = torch.empty(q.shape[0], q.shape[2], q.shape[1], device=q.device, dtype=torch.float32)
    else:
        o = torch.zeros_like(q)
        lse = torch.zeros(cu_q.numel() - 1, q.shape[1], int(max_sq), device=q.device, dtype=torch.float32)
    p = make_fwd_params(q, k, v, o, lse, causal, cu_q, cu_k, max_sq, max_sk)
    rc = lib.fa_b200_fwd(ctypes.byref(p), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != FA_OK:
        raise RuntimeError(f"fa_b200_fwd rc={rc}: {lib.fa_b200_last_error().decode()}")
    return o, lse


def bwd(q, k, v, o, lse, dout, causal, cu_q=None, cu_k=None, max_sq=None, max_sk=None, use_workspace=True):
    import torch
    lib = load()
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    dsum = torch.zeros_like(lse)
    p = BwdParams()
    p.fwd = make_fwd_params(q, k, v, o, lse, causal, cu_q, cu_k, max_sq, max_sk)
    p.dout, p.dq, p.dk, p.dv, p.dsum = dout.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dsum.data_ptr()
    nbytes = lib.fa_b200_bwd_workspace_bytes(ctypes.byref(p.fwd)) if use_workspace else 0
    ws = torch.empty(max(int(nbytes), 1), device=q.device, dtype=torch.uint8)
    p.workspace = ws.data_ptr() if nbytes > 0 else None
    rc = lib.fa_b200_bwd(ctypes.byref(p), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != FA_OK:
        raise RuntimeError(f"fa_b200_bwd rc={rc}: {lib.fa_b200_last_error().decode()}")
    return dq, dk, dv
