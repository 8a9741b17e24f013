import os, sys, time
sys.path.insert(0, "/root/repo/flash-attention-turing_b200")
import torch, flash_attn_turing as fat
torch.manual_seed(0)
b, s = 4, 4096
q = torch.randn(b, s, 32, 128, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
for i in range(30):
    t0 = time.time()
    o, l = fat.fwd(q, k, v, False)
    torch.cuda.synchronize()
    print("launch", i, "ok", round((time.time() - t0) * 1e3, 3), "ms", flush=True)
for i in range(5):
    t0 = time.time()
    for _ in range(10): fat.fwd(q, k, v, False)
    torch.cuda.synchronize()
    print("batch", i, "ok", round((time.time() - t0) * 1e2, 3), "ms/launch", flush=True)
