"""CPU (gloo, world_size 2 and 3) test of the multi-GPU host logic: batch partitioning, scatter, gather.  The per-rank
compute is the CPU oracle here (tests may use it as the checker); on a GPU box it is flash_attn_turing.fwd."""
import importlib.util
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_sharded():
    # load sharded.py without importing the package __init__ (which needs the CUDA extension + a GPU-side libcudart)
    path = os.path.join(ROOT, "flash-attention-turing_b200", "flash_attn_turing", "sharded.py")
    spec = importlib.util.spec_from_file_location("fat_sharded", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _oracle_fwd(q, k, v, causal):
    sys.path.insert(0, ROOT)
    from oracle import oracle
    o, l = oracle.attention_fwd(q.numpy(), k.numpy(), v.numpy(), causal)
    return torch.from_numpy(o), torch.from_numpy(l)


def _worker(rank, world, port, batch, causal, result_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = _load_sharded()
    torch.manual_seed(0)
    shapes = ((batch, 33, 4, 64), (batch, 47, 2, 64))
    q = k = v = None
    if rank == 0:
        q, k, v = torch.randn(shapes[0]), torch.randn(shapes[1]), torch.randn(shapes[1])
    o, l = sh.fwd_sharded(q, k, v, causal, _oracle_fwd, shapes=shapes, dtype=torch.float32, device="cpu")
    if rank == 0:
        o_ref, l_ref = _oracle_fwd(q, k, v, causal)
        np.save(result_path, np.array([float((o - o_ref).abs().max()), float((l - l_ref).abs().max())]))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_host_pipeline_chunk_ranges():
    from flash_attn_turing.hostio import chunk_ranges
    assert chunk_ranges(4, 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]
    assert chunk_ranges(5, 2) == [(0, 3), (3, 5)]
    assert chunk_ranges(2, 8) == [(0, 1), (1, 2)]
    for b in range(0, 9):
        for n in range(1, 6):
            r = chunk_ranges(b, n)
            assert sum(e - s for s, e in r) == b and all(e > s for s, e in r)


def test_host_pipeline_chunk_plan():
    """batch ranges first; beyond `batch` chunks the KV heads of every batch are split into equal groups"""
    from flash_attn_turing.hostio import plan_chunks
    assert plan_chunks(4, 32, 4) == [(0, 1, 0, 32), (1, 2, 0, 32), (2, 3, 0, 32), (3, 4, 0, 32)]
    assert plan_chunks(4, 32, None) == [(b, b + 1, 8 * g, 8 * g + 8) for b in range(4) for g in range(4)]
    assert plan_chunks(5, 2, None) == [(b, b + 1, g, g + 1) for b in range(5) for g in range(2)]
    assert plan_chunks(5, 2, 2) == [(0, 3, 0, 2), (3, 5, 0, 2)]
    assert plan_chunks(2, 6, 9) == [(b, b + 1, 2 * g, 2 * g + 2) for b in range(2) for g in range(3)]   # 4 per batch -> 3 divides 6
    assert plan_chunks(3, 1, 16) == [(0, 1, 0, 1), (1, 2, 0, 1), (2, 3, 0, 1)]
    assert plan_chunks(0, 4, None) == []
    for b in range(1, 7):
        for hk in (1, 2, 3, 8):
            for n in (None, 1, 2, 5, 16, 64):
                cover = set()
                for b0, b1, g0, g1 in plan_chunks(b, hk, n):
                    for bi in range(b0, b1):
                        for g in range(g0, g1):
                            assert (bi, g) not in cover
                            cover.add((bi, g))
                assert len(cover) == b * hk


def test_shard_ranges():
    sh = _load_sharded()
    assert sh.shard_ranges(256, 8) == [(32 * i, 32 * i + 32) for i in range(8)]
    assert sh.shard_ranges(5, 3) == [(0, 2), (2, 4), (4, 5)]
    assert sh.shard_ranges(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    for b in range(0, 20):
        for w in range(1, 9):
            r = sh.shard_ranges(b, w)
            assert r[0][0] == 0 and r[-1][1] == b and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


@pytest.mark.parametrize("world,batch,causal", [(2, 4, False), (2, 5, True), (3, 2, False)])
def test_batch_shard_scatter_compute_gather(tmp_path, world, batch, causal):
    """identical to the single-rank result, including uneven splits and ranks that receive no batch at all"""
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, _free_port(), batch, causal, out), nprocs=world, join=True)
    err = np.load(out)
    assert err[0] == 0.0 and err[1] == 0.0, err
