"""Batch-shard runner: the (batch x head) attention problems are independent, so N GPUs each take a contiguous slab of
the batch dimension and run the single-GPU kernels — no collective on the data path.  torch.distributed
is used only to move slabs when the caller holds the whole batch on one rank (scatter Q/K/V, gather O) and lives outside
any timed region.  The reference has no multi-GPU support at all (its int32 offsets overflow on the unsharded config 5,
block_info.h:15-21).

The compute callable is injected so that the host logic can be exercised on CPU (gloo) in tests; in production it is
flash_attn_turing.fwd.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_ranges(batch: int, world_size: int) -> List[Tuple[int, int]]:
    """contiguous [start, end) batch ranges per rank; the first `batch % world_size` ranks get one extra"""
    base, extra = divmod(batch, world_size)
    out, s = [], 0
    for r in range(world_size):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def _run_p2p(ops) -> None:
    """one batched group of sends / receives: NCCL then moves the slabs of all peers concurrently (unbatched isend / recv
    calls are serialised per process group — measured 267 GB/s out of rank 0 on an NVSwitch box, profiles/r02b_n8.log)"""
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def _scatter_ops(x: Optional[torch.Tensor], shape, dtype, device, src: int, group, keep: list):
    """P2P ops that scatter x [b, ...] from rank `src`; returns (ops, own slab)"""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = shard_ranges(shape[0], world)
    s, e = ranges[rank]
    out = torch.empty((e - s,) + tuple(shape[1:]), dtype=dtype, device=device)
    ops = []
    if rank == src:
        for r, (rs, re) in enumerate(ranges):
            if r == src:
                out.copy_(x[rs:re])
            elif re > rs:
                slab = x[rs:re].contiguous()          # a batch slice of a contiguous tensor: no copy
                keep.append(slab)
                ops.append(dist.P2POp(dist.isend, slab, r, group))
    elif e > s:
        ops.append(dist.P2POp(dist.irecv, out, src, group))
    return ops, out


def _gather_ops(x_local: torch.Tensor, batch: int, dst: int, group, keep: list):
    """P2P ops that gather the slabs on rank `dst`; returns (ops, full tensor or None)"""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = shard_ranges(batch, world)
    ops = []
    if rank == dst:
        full = torch.empty((batch,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        for r, (rs, re) in enumerate(ranges):
            if r == dst:
                full[rs:re].copy_(x_local)
            elif re > rs:
                ops.append(dist.P2POp(dist.irecv, full[rs:re], r, group))
        return ops, full
    if x_local.shape[0] > 0:
        slab = x_local.contiguous()
        keep.append(slab)
        ops.append(dist.P2POp(dist.isend, slab, dst, group))
    return ops, None


def scatter_batch(x: Optional[torch.Tensor], shape, dtype, device, src: int = 0, group=None) -> torch.Tensor:
    """rank `src` holds x [b, ...]; every rank returns its slab [b_r, ...] (contiguous slices: zero repacking)"""
    keep: list = []
    ops, out = _scatter_ops(x, shape, dtype, device, src, group, keep)
    _run_p2p(ops)
    return out


def gather_batch(x_local: torch.Tensor, batch: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """inverse of scatter_batch: rank `dst` returns the full [b, ...] tensor, others None"""
    keep: list = []
    ops, full = _gather_ops(x_local, batch, dst, group, keep)
    _run_p2p(ops)
    return full


def fwd_sharded(q, k, v, is_causal: bool, fwd_fn: Callable, shapes=None, dtype=None, device=None, src: int = 0, group=None):
    """Full-batch forward across the ranks of `group`.  On rank `src` q,k,v are the full tensors (None elsewhere, then
    `shapes` = (q.shape, k.shape), dtype and device must be given).  Returns (o, l) on rank `src`, (None, None) elsewhere.
    Q, K and V leave rank `src` in ONE batched group of sends, O and LSE come back in one group of receives."""
    rank = dist.get_rank(group)
    if rank == src:
        shapes, dtype, device = (tuple(q.shape), tuple(k.shape)), q.dtype, q.device
    keep: list = []
    ops_q, ql = _scatter_ops(q, shapes[0], dtype, device, src, group, keep)
    ops_k, kl = _scatter_ops(k, shapes[1], dtype, device, src, group, keep)
    ops_v, vl = _scatter_ops(v, shapes[1], dtype, device, src, group, keep)
    _run_p2p(ops_q + ops_k + ops_v)
    if ql.shape[0] > 0:
        ol, ll = fwd_fn(ql, kl, vl, is_causal)
    else:
        ol = ql
        ll = torch.empty((0, shapes[0][2], shapes[0][1]), dtype=torch.float32, device=device)
    keep = []
    ops_o, o = _gather_ops(ol, shapes[0][0], src, group, keep)
    ops_l, l = _gather_ops(ll, shapes[0][0], src, group, keep)
    _run_p2p(ops_o + ops_l)
    return o, l
