"""2-rank NCCL test of the batch-shard I/O path (north_star: "NCCL over NVLink only to scatter Q/K/V and gather O"):
flash_attn_turing.sharded.fwd_sharded with the `nccl` backend must be bit-identical to the single-GPU call."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, time
sys.path.insert(0, os.path.join(sys.argv[1], "flash-attention-turing_b200"))
import torch, torch.distributed as dist
import flash_attn_turing as fat
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
for (b, s, h, hk, d, causal) in [(5, 700, 4, 2, 128, True), (8, 1024, 8, 8, 128, False), (3, 333, 6, 3, 64, True)]:
    shapes = ((b, s, h, d), (b, s, hk, d))
    q = k = v = None
    if rank == 0:
        torch.manual_seed(b * s)
        q, k, v = (torch.randn(*shapes[0], device=dev, dtype=torch.bfloat16), torch.randn(*shapes[1], device=dev, dtype=torch.bfloat16),
                   torch.randn(*shapes[1], device=dev, dtype=torch.bfloat16))
    o, l = fat.sharded.fwd_sharded(q, k, v, causal, fat.fwd, shapes=shapes, dtype=torch.bfloat16, device=dev)
    torch.cuda.synchronize()
    if rank == 0:
        o1, l1 = fat.fwd(q, k, v, causal)
        assert torch.equal(o, o1) and torch.equal(l, l1), (b, s, h, hk, d, causal)
        print("SHARD OK", b, s, h, hk, d, causal, flush=True)
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fwd_sharded_over_nccl_two_ranks(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script), ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count("SHARD OK") == 3, r.stdout[-3000:] + r.stderr[-3000:]
