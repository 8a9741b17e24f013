"""On the GPU box: time the reference's own kernels (rebuilt for sm_100a, fp16, baseline/_ref) beside ours (fp16 and
bf16) and torch SDPA on the BASELINE shapes.  Prints markdown rows for BASELINE.md §5."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import torch, torch.nn.functional as F
import flash_attn_turing as ours
try:
    import flash_attn_turing_ref as ref
except Exception as e:  # noqa
    ref = None; print("reference .so not available:", e)

def timeit(fn, n):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

shapes = [("C2", 4, 4096, False), ("C3", 4, 8192, True), ("C4", 4, 16384, False)]
print("| config | impl | dtype | fwd ms | fwd TFLOP/s | bwd ms | bwd TFLOP/s |")
print("|---|---|---|---|---|---|---|")
for nm, b, s, causal in shapes:
    fl = 4 * b * 32 * s * s * 128 * (0.5 if causal else 1.0)
    n = 10 if s <= 8192 else 3
    for dt in (torch.float16, torch.bfloat16):
        torch.manual_seed(0)
        q = torch.randn(b, s, 32, 128, device="cuda", dtype=dt); k = torch.randn_like(q); v = torch.randn_like(q); do = torch.randn_like(q)
        impls = [("ours", ours)] + ([("reference kernels (sm_100a rebuild, HMMA.1688)", ref)] if (ref is not None and dt == torch.float16) else [])
        for name, m in impls:
            o, l = m.fwd(q, k, v, causal)
            tf = timeit(lambda: m.fwd(q, k, v, causal), n)
            tb = timeit(lambda: m.bwd(q, k, v, o, l, do, causal), max(2, n // 2))
            print(f"| {nm} | {name} | {str(dt)[6:]} | {tf:.3f} | {fl/tf/1e9:.0f} | {tb:.3f} | {2.5*fl/tb/1e9:.0f} |", flush=True)
        qt, kt, vt = [t.transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v)]
        tf = timeit(lambda: F.scaled_dot_product_attention(qt, kt, vt, is_causal=causal), n)
        ot = F.scaled_dot_product_attention(qt, kt, vt, is_causal=causal)
        tb = timeit(lambda: torch.autograd.grad(ot, (qt, kt, vt), do.transpose(1, 2), retain_graph=True), max(2, n // 2))
        print(f"| {nm} | torch SDPA (fused backend) | {str(dt)[6:]} | {tf:.3f} | {fl/tf/1e9:.0f} | {tb:.3f} | {2.5*fl/tb/1e9:.0f} |", flush=True)
