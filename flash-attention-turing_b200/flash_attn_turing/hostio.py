"""Forward attention on HOST-resident tensors: `fwd_host(q, k, v, is_causal)`.

The (batch x head) problems are independent, so a batch that lives in (pinned) host memory is cut into batch chunks and
streamed through the GPU: while chunk c is being computed, chunk c+1 is on its way up the PCIe link and the result of
chunk c-1 is on its way down (three CUDA streams; H2D and D2H use the two directions of the link concurrently).  The
kernel is the same single launch per chunk as `fwd`; nothing here touches the math.  At BASELINE config 2 (b4 s4096 h32
d128 bf16: 403 MB up, 136 MB down) the serial copy -> compute -> copy sequence costs ~10.5 ms, the pipeline ~7.5 ms —
the floor is the upstream copy alone (403 MB at ~55 GB/s).

The reference has no host-side entry point (its callers always pass CUDA tensors, test_flash_attn.py:352-380); this is
the end-to-end path bench.py reports under "e2e".
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

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

    def _ensure(self, key, chunk_b, q, k):
        if self._key == key:
            return
        self._key = key
        self._bufs = []
        for _ in range(self.slots):
            self._bufs.append(tuple(torch.empty((chunk_b,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
                                    for t in (q, k, k)))
        self._free = [None] * self.slots

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
        b, sq, h, d = q.shape
        if out is None:
            out = torch.empty(q.shape, dtype=q.dtype).pin_memory()
        if lse is None:
            lse = torch.empty((b, h, sq), dtype=torch.float32).pin_memory()
        if b == 0:
            return out, lse
        ranges = chunk_ranges(b, b if chunks is None else chunks)
        chunk_b = max(e - s for s, e in ranges)
        self._ensure((tuple(q.shape), tuple(k.shape), q.dtype, chunk_b), chunk_b, q, k)
        cur = torch.cuda.current_stream(self.device)
        self.s_h2d.wait_stream(cur)          # the caller's earlier work on these buffers is ordered before the copies
        done = None
        for c, (s, e) in enumerate(ranges):
            slot = c % self.slots
            dq, dk, dv = (t[: e - s] for t in self._bufs[slot])
            with torch.cuda.stream(self.s_h2d):
                if self._free[slot] is not None:
                    self.s_h2d.wait_event(self._free[slot])
                dq.copy_(q[s:e], non_blocking=True)
                dk.copy_(k[s:e], non_blocking=True)
                dv.copy_(v[s:e], non_blocking=True)
                up = torch.cuda.Event()
                up.record(self.s_h2d)
            cur.wait_event(up)
            if copy_only:
                if self._scratch is None or self._scratch[0].shape[0] < e - s or self._scratch[0].shape[1:] != dq.shape[1:]:
                    self._scratch = (torch.empty_like(self._bufs[0][0]), torch.empty((chunk_b, h, sq), dtype=torch.float32, device=self.device))
                o_c, l_c = self._scratch[0][: e - s], self._scratch[1][: e - s]
            else:
                o_c, l_c = self.fwd_fn(dq, dk, dv, is_causal)      # one kernel launch on the current stream
            ev = torch.cuda.Event()
            ev.record(cur)
            self._free[slot] = ev
            o_c.record_stream(self.s_d2h)
            l_c.record_stream(self.s_d2h)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(ev)
                out[s:e].copy_(o_c, non_blocking=True)
                lse[s:e].copy_(l_c, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.s_d2h)
        self.last_chunks = len(ranges)
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
