"""flash_attn_turing — the reference's operator surface, backed by hand-written sm_100a (B200) kernels.

Drop-in for ssiu/flash-attention-turing's extension module (/root/reference/csrc/flash_attn/flash_api.cpp:471-476):

    from flash_attn_turing import fwd, bwd, varlen_fwd, varlen_bwd

    fwd(q, k, v, is_causal)                               -> [o, l]
    bwd(q, k, v, o, l, dout, is_causal)                   -> [dq, dk, dv]
    varlen_fwd(q, k, v, cu_q, cu_k, max_sq, max_sk, is_causal)             -> [out, l]
    varlen_bwd(q, k, v, out, l, dout, cu_q, cu_k, max_sq, max_sk, is_causal) -> [dq, dk, dv]

plus `flash_attn_func`, the name the reference's README documents (/root/reference/README.md:28-47) and
utils/scratch_debug.py:4,14 still imports.

There is NO fallback: if the compiled extension (`_C`, linked against libfa_b200.so) is missing, importing this
package raises.
"""
from __future__ import annotations

import os as _os

import torch as _torch

try:
    from . import _C  # noqa: F401  (pybind11 extension built by setup.py / __graft_entry__.build())
except ImportError as _e:  # pragma: no cover - loud by design
    raise ImportError(
        "flash_attn_turing: the compiled extension flash_attn_turing._C (and libfa_b200.so) is not built. "
        "Run `python setup.py build_ext --inplace` (or `python -c 'import __graft_entry__ as g; g.build()'`) "
        "at the repo root. There is no CPU / PyTorch fallback path."
    ) from _e

fwd = _C.fwd
bwd = _C.bwd
varlen_fwd = _C.varlen_fwd
varlen_bwd = _C.varlen_bwd
last_launch_count = _C.last_launch_count

LIB_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "libfa_b200.so")


class _FlashAttnFunc(_torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal):
        o, l = fwd(q, k, v, causal)
        ctx.save_for_backward(q, k, v, o, l)
        ctx.causal = causal
        return o

    @staticmethod
    def backward(ctx, dout):
        q, k, v, o, l = ctx.saved_tensors
        dq, dk, dv = bwd(q, k, v, o, l, dout.contiguous(), ctx.causal)
        return dq, dk, dv, None


def flash_attn_func(q, k, v, *legacy_dims, causal: bool = False):
    """README-era entry point: ``flash_attn_func(q, k, v, batch_size, seq_len, num_heads, head_dim) -> out``.

    q: [batch, seqlen_q, heads, head_dim]; k, v: [batch, seqlen_k, heads_k, head_dim]; fp16 or bf16, CUDA,
    contiguous.  The four legacy integer arguments are accepted and ignored (shapes are read from the tensors,
    as the current reference code does, flash_api.cpp:166-176).  Differentiable (autograd over fwd/bwd).
    """
    if len(legacy_dims) not in (0, 4):
        raise TypeError("flash_attn_func(q, k, v[, batch_size, seq_len, num_heads, head_dim], causal=False)")
    return _FlashAttnFunc.apply(q, k, v, bool(causal))


from . import sharded  # noqa: E402,F401  (batch-shard runner for multi-GPU boxes)
from .hostio import HostForward, fwd_host  # noqa: E402,F401  (host-resident tensors: chunked copy/compute pipeline)

__all__ = ["fwd", "bwd", "varlen_fwd", "varlen_bwd", "flash_attn_func", "last_launch_count", "LIB_PATH", "sharded",
           "fwd_host", "HostForward"]
