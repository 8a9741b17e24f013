"""ctypes binding of the CPU oracle (oracle/attn_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product package (flash_attn_turing) never does.  See the header of attn_oracle.c for what is restated and
how the oracle is pinned (tests/golden/*.npz generated from the reference's own vanilla_attention_ref).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")


class _Dims(ctypes.Structure):
    _fields_ = [("b", ctypes.c_int64), ("sq", ctypes.c_int64), ("sk", ctypes.c_int64), ("h", ctypes.c_int64),
                ("h_k", ctypes.c_int64), ("d", ctypes.c_int64), ("cu_q", ctypes.c_void_p), ("cu_k", ctypes.c_void_p),
                ("causal", ctypes.c_int32)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "attn_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _dims(q, k, causal, cu_q, cu_k, max_sq, max_sk):
    if cu_q is None:
        b, sq, h, d = q.shape
        sk, h_k = k.shape[1], k.shape[2]
        keep = ()
        dims = _Dims(b, sq, sk, h, h_k, d, None, None, int(bool(causal)))
    else:
        cu_q = np.ascontiguousarray(np.asarray(cu_q, dtype=np.int32))
        cu_k = np.ascontiguousarray(np.asarray(cu_k, dtype=np.int32))
        _, h, d = q.shape
        h_k = k.shape[1]
        keep = (cu_q, cu_k)
        dims = _Dims(len(cu_q) - 1, int(max_sq), int(max_sk), h, h_k, d, cu_q.ctypes.data, cu_k.ctypes.data,
                     int(bool(causal)))
    return dims, keep


def attention_fwd(q, k, v, causal=False, cu_q=None, cu_k=None, max_sq=None, max_sk=None, fast=False):
    """q [b,sq,h,d] (or packed [total,h,d] with cu_q/cu_k) float arrays -> (o float32 like q, lse float32 [b,h,sq])."""
    q, k, v = _f32(q), _f32(k), _f32(v)
    dims, keep = _dims(q, k, causal, cu_q, cu_k, max_sq, max_sk)
    o = np.zeros_like(q)
    lse = np.zeros((dims.b, dims.h, dims.sq), dtype=np.float32)
    fn = _load().oracle_fwd_f32 if fast else _load().oracle_fwd
    fn(ctypes.byref(dims), _fp(q), _fp(k), _fp(v), _fp(o), _fp(lse))
    del keep
    return o, lse


def attention_bwd(q, k, v, o, lse, dout, causal=False, cu_q=None, cu_k=None, max_sq=None, max_sk=None):
    """-> (dq, dk, dv) float32, shapes of q, k, v."""
    q, k, v, o, lse, dout = _f32(q), _f32(k), _f32(v), _f32(o), _f32(lse), _f32(dout)
    dims, keep = _dims(q, k, causal, cu_q, cu_k, max_sq, max_sk)
    dq, dk, dv = np.zeros_like(q), np.zeros_like(k), np.zeros_like(v)
    _load().oracle_bwd(ctypes.byref(dims), _fp(q), _fp(k), _fp(v), _fp(o), _fp(lse), _fp(dout), _fp(dq), _fp(dk), _fp(dv))
    del keep
    return dq, dk, dv
