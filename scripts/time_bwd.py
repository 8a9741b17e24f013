"""backward timing (device-resident, CUDA events): usage time_bwd.py [name ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200"))
import torch, flash_attn_turing as fat
import torch.nn.functional as F
shapes = {"C2": (4, 4096, False), "C3": (4, 8192, True), "C4": (4, 16384, False), "S2k": (8, 2048, False), "C2c": (4, 4096, True)}
for nm in (sys.argv[1:] or ["C2", "C2c", "C4"]):
    b, s, causal = shapes[nm]
    torch.manual_seed(0)
    q = torch.randn(b, s, 32, 128, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q); do = torch.randn_like(q)
    o, l = fat.fwd(q, k, v, causal)
    for _ in range(2): g = fat.bwd(q, k, v, o, l, do, causal)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10 if s <= 8192 else 4
    e0.record()
    for _ in range(n): fat.bwd(q, k, v, o, l, do, causal)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.5 * 4 * b * 32 * s * s * 128 * (0.5 if causal else 1.0)
    # reference: torch SDPA backward on one (b) slice for error, and its timing on the full shape
    qt, kt, vt = [t.transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v)]
    ot = F.scaled_dot_product_attention(qt, kt, vt, is_causal=causal)
    gt = torch.autograd.grad(ot, (qt, kt, vt), do.transpose(1, 2), retain_graph=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n): torch.autograd.grad(ot, (qt, kt, vt), do.transpose(1, 2), retain_graph=True)
    e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / n
    errs = [(a.float() - b_.transpose(1, 2).float()).abs().max().item() for a, b_ in zip(g, gt)]
    print(f"TIMING bwd {nm} b{b} s{s} causal={causal}: ours {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s | torch SDPA bwd {ms_t:.3f} ms {fl / ms_t / 1e9:.1f} TFLOP/s | max|diff| dq,dk,dv {errs}", flush=True)
