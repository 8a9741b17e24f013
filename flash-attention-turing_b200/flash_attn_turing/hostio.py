"""Forward attention on HOST-resident tensors: `fwd_host(q, k, v, is_causal)`.

The (batch x head) problems are independent, so a batch that lives in (pinned) host memory is cut into chunks and
streamed through the GPU: while chunk c is being computed, chunk c+1 is on its way up the PCIe link and the result of
chunk c-1 is on its way down (three CUDA streams; H2D and D2H use the two directions of the link concurrently).  The
kernel is the same single launch per chunk as `fwd`; nothing here touches the math.

Chunks are batch ranges and, when there are fewer batches than pipeline stages want, KV-head groups inside one batch
(a head group of a [b, s, h, d] tensor is a pitched 2-D region: rows of `heads * d` elements, pitch `h * d` — moved with
cudaMemcpy2DAsync, still one DMA per tensor).  At BASELINE config 2 (b4 s4096 h32 d128 bf16: 403 MB up, 136 MB down)
the serial copy -> compute -> copy sequence costs ~10.5 ms, four per-batch chunks 8.5 ms, sixteen (batch, 8-head) chunks
shave most of the remaining fill/drain; the floor is the upstream copy alone (403 MB at ~55 GB/s = 7.3 ms).

The reference has no host-side entry point (its callers always pass CUDA tensors, test_flash_attn.py:352-380); this is
the end-to-end path bench.py reports under "e2e".
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Optional, Tuple

import torch


def chunk_ranges(batch: int, n_chunks: int):
    """contiguous [start, end) batch ranges, sizes differing by at most one, empty chunks dropped"""
    n_chunks = max(1, min(int(n_chunks), batch)) if batch > 0 else 1
    base, extra = divmod(batch, n_chunks)
    out, s = [], 0
    for c in range(n_chunks):
        e = s + base + (1 if c < extra else 0)
        if e > s:
            out.append((s, e))
        s = e
    return out


def plan_chunks(batch: int, heads_k: int, n_chunks: Optional[int]) -> List[Tuple[int, int, int, int]]:
    """-> [(b0, b1, kg0, kg1)]: batch range x KV-head range per chunk.  Up to `batch` chunks are whole batches; beyond that
    every batch is one chunk row and its KV heads are split into equal groups (group count = the largest divisor of
    heads_k that keeps the total at or below n_chunks).  n_chunks=None: aim for 16 pipeline stages."""
    if batch <= 0:
        return []
    want = 16 if n_chunks is None else max(1, int(n_chunks))
    if want <= batch or heads_k <= 1:
        return [(s, e, 0, heads_k) for s, e in chunk_ranges(batch, want)]
    per_batch = max(1, want // batch)
    groups = max(g for g in range(1, heads_k + 1) if heads_k % g == 0 and g <= per_batch)
    hg = heads_k // groups
    return [(b0, b0 + 1, g * hg, (g + 1) * hg) for b0 in range(batch) for g in range(groups)]


_cudart = None


def _memcpy2d_async(dst_ptr, dpitch, src_ptr, spitch, width, height, kind, stream):
    """cudaMemcpy2DAsync through the CUDA runtime already loaded by torch (torch exposes no 2-D copy)"""
    global _cudart
    if _cudart is None:
        _cudart = ctypes.CDLL("libcudart.so.12")
        _cudart.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                              ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        _cudart.cudaMemcpy2DAsync.restype = ctypes.c_int
    rc = _cudart.cudaMemcpy2DAsync(dst_ptr, dpitch, src_ptr, spitch, width, height, kind, stream)
    if rc != 0:
        raise RuntimeError(f"cudaMemcpy2DAsync failed with cudaError {rc}")


class HostForward:
    """Reusable pipeline state (streams, device staging slots) for repeated `fwd_host` calls of one shape."""

    def __init__(self, fwd_fn: Optional[Callable] = None, device=None, slots: int = 3):
        if fwd_fn is None:
            from . import fwd as fwd_fn  # the sm_100a kernel; no fallback
        self.fwd_fn = fwd_fn
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.slots = max(2, int(slots))
        self.s_h2d = torch.cuda.Stream(device=self.device)
        self.s_d2h = torch.cuda.Stream(device=self.device)
        self._key = None
        self._bufs = []
        self._free = []      # per slot: event recorded on the compute stream once the slot's inputs have been consumed
        self._scratch = None  # copy_only: device o / lse stand-ins for the D2H leg
        self.last_chunks = 0  # kernel launches (= chunks) of the last call

    def _ensure(self, key, cb, q, k, hq, hk):
        if self._key == key:
            return
        self._key = key
        self._bufs = []
        for _ in range(self.slots):
            self._bufs.append((torch.empty((cb, q.shape[1], hq, q.shape[3]), dtype=q.dtype, device=self.device),
                               torch.empty((cb, k.shape[1], hk, k.shape[3]), dtype=k.dtype, device=self.device),
                               torch.empty((cb, k.shape[1], hk, k.shape[3]), dtype=k.dtype, device=self.device)))
        self._free = [None] * self.slots
        self._scratch = None

    def __call__(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, is_causal: bool,
                 out: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None, chunks: Optional[int] = None,
                 sync: bool = True, copy_only: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """q [b,sq,h,d], k/v [b,sk,h_k,d] on the host (pinned memory makes the copies asynchronous).  Returns (o, lse)
        as host tensors (`out` / `lse` if given, else new pinned tensors).  With sync=False the caller must synchronize
        the current stream before reading them.  copy_only=True moves the same bytes in the same chunks but launches no
        kernel (the outputs are garbage): bench.py uses it to measure the copy floor of this host."""
        if q.is_cuda or k.is_cuda or v.is_cuda:
            raise ValueError("fwd_host expects host tensors; use flash_attn_turing.fwd for CUDA tensors")
        if q.dim() != 4 or k.shape != v.shape or k.dim() != 4 or q.shape[0] != k.shape[0]:
            raise ValueError("q must be [b,sq,h,d] and k, v [b,sk,h_k,d] with the same batch")
        if not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
            raise ValueError("q, k, v must be contiguous")
        b, sq, h, d = q.shape
        sk, h_k = k.shape[1], k.shape[2]
        if h_k <= 0 or h % h_k != 0:
            raise ValueError("num_heads_q must be divisible by num_heads_k for GQA/MQA")
        ratio = h // h_k
        if out is None:
            out = torch.empty(q.shape, dtype=q.dtype).pin_memory()
        if lse is None:
            lse = torch.empty((b, h, sq), dtype=torch.float32).pin_memory()
        if b == 0:
            return out, lse
        plan = plan_chunks(b, h_k, chunks)
        cb = max(b1 - b0 for b0, b1, _, _ in plan)
        hk_c = max(g1 - g0 for _, _, g0, g1 in plan)
        self._ensure((tuple(q.shape), tuple(k.shape), q.dtype, cb, hk_c), cb, q, k, hk_c * ratio, hk_c)
        es = q.element_size()
        cur = torch.cuda.current_stream(self.device)
        self.s_h2d.wait_stream(cur)          # the caller's earlier work on these buffers is ordered before the copies
        done = None
        for c, (b0, b1, g0, g1) in enumerate(plan):
            slot = c % self.slots
            nb, hk_n = b1 - b0, g1 - g0
            hq_n, hq0 = hk_n * ratio, g0 * ratio
            whole = hk_n == h_k
            bq, bk, bv = self._bufs[slot]
            if whole:
                dq, dk, dv = bq[:nb], bk[:nb], bv[:nb]
            else:   # one batch, a head group: contiguous [1, s, heads, d] views of the slot's storage
                dq = bq.view(-1)[: sq * hq_n * d].view(1, sq, hq_n, d)
                dk = bk.view(-1)[: sk * hk_n * d].view(1, sk, hk_n, d)
                dv = bv.view(-1)[: sk * hk_n * d].view(1, sk, hk_n, d)
            with torch.cuda.stream(self.s_h2d):
                if self._free[slot] is not None:
                    self.s_h2d.wait_event(self._free[slot])
                if whole:
                    dq.copy_(q[b0:b1], non_blocking=True)
                    dk.copy_(k[b0:b1], non_blocking=True)
                    dv.copy_(v[b0:b1], non_blocking=True)
                else:
                    st = self.s_h2d.cuda_stream
                    _memcpy2d_async(dq.data_ptr(), hq_n * d * es, q[b0, 0, hq0].data_ptr(), h * d * es, hq_n * d * es, sq, 1, st)
                    _memcpy2d_async(dk.data_ptr(), hk_n * d * es, k[b0, 0, g0].data_ptr(), h_k * d * es, hk_n * d * es, sk, 1, st)
                    _memcpy2d_async(dv.data_ptr(), hk_n * d * es, v[b0, 0, g0].data_ptr(), h_k * d * es, hk_n * d * es, sk, 1, st)
                up = torch.cuda.Event()
                up.record(self.s_h2d)
            cur.wait_event(up)
            if copy_only:
                if self._scratch is None:
                    self._scratch = (torch.empty_like(bq), torch.empty((cb, hk_c * ratio, sq), dtype=torch.float32, device=self.device))
                o_c = self._scratch[0].view(-1)[: dq.numel()].view(dq.shape)
                l_c = self._scratch[1].view(-1)[: nb * hq_n * sq].view(nb, hq_n, sq)
            else:
                o_c, l_c = self.fwd_fn(dq, dk, dv, is_causal)      # one kernel launch on the current stream
            ev = torch.cuda.Event()
            ev.record(cur)
            self._free[slot] = ev
            o_c.record_stream(self.s_d2h)
            l_c.record_stream(self.s_d2h)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(ev)
                if whole:
                    out[b0:b1].copy_(o_c, non_blocking=True)
                    lse[b0:b1].copy_(l_c, non_blocking=True)
                else:
                    _memcpy2d_async(out[b0, 0, hq0].data_ptr(), h * d * es, o_c.data_ptr(), hq_n * d * es, hq_n * d * es, sq, 2,
                                    self.s_d2h.cuda_stream)
                    lse[b0, hq0:hq0 + hq_n].copy_(l_c[0], non_blocking=True)   # [heads, sq] rows of one batch: contiguous
                done = torch.cuda.Event()
                done.record(self.s_d2h)
        self.last_chunks = len(plan)
        cur.wait_event(done)                 # stream order: work queued after this call sees the results on the host
        if sync:
            cur.synchronize()
        return out, lse


_default: dict = {}


def fwd_host(q, k, v, is_causal: bool, out=None, lse=None, chunks: Optional[int] = None, sync: bool = True):
    """module-level convenience wrapper around a cached HostForward for the current device"""
    dev = torch.cuda.current_device()
    hf = _default.get(dev)
    if hf is None:
        hf = _default[dev] = HostForward()
    return hf(q, k, v, is_causal, out=out, lse=lse, chunks=chunks, sync=sync)
