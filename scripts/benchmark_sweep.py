"""benchmark.sh-compatible sweep (reference: /root/reference/benchmark.sh:17-45): for every (seqlen, head_dim, causal) of
the reference's grid run fwd+bwd at batch 4, 16 heads and print a text table (the reference pipes ncu CSVs into
matplotlib, which this image does not have; kernel times here come from CUDA events around each operator call).

    python scripts/benchmark_sweep.py [--dtype fp16|bf16] [--quick]      (GPU box)
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200"))
import torch
import torch.nn.functional as F
import flash_attn_turing as fat

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--quick", action="store_true")
ap.add_argument("--hdim", type=int, default=0, help="only this head_dim (0 = both, as benchmark.sh:20)")
args = ap.parse_args()
dt = torch.float16 if args.dtype == "fp16" else torch.bfloat16
batch_size, num_heads = 4, 16                                          # benchmark.sh:17-19
seqlens = [512, 1024, 2048, 4096, 8192, 16384, 500, 1000, 2000, 4000, 8000, 16000]   # benchmark.sh:21
if args.quick:
    seqlens = [512, 4096, 4000]


def timeit(fn, n):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ref = None
try:   # the reference's own kernels rebuilt for sm_100a (fp16 only), when staged by baseline/build_ref.sh
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import flash_attn_turing_ref as ref
except Exception:
    ref = None
print(f"benchmark.sh grid (batch {batch_size}, heads {num_heads}), dtype {args.dtype}; CUDA events, device-resident tensors; TFLOP/s algorithmic (bwd = 2.5 x fwd)")
print(f"| seqlen | hdim | causal | flash_fwd_kernel ms | TFLOP/s | flash_bwd (dot+dq+dk_dv) ms | TFLOP/s | torch SDPA fwd ms | TFLOP/s | torch SDPA bwd ms | TFLOP/s | reference kernels fwd ms | bwd ms |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for s in seqlens:
    for d in ((args.hdim,) if args.hdim else (64, 128)):
        for causal in (False, True):
            torch.manual_seed(0)
            q = torch.randn(batch_size, s, num_heads, d, device="cuda", dtype=dt); k = torch.randn_like(q); v = torch.randn_like(q); do = torch.randn_like(q)
            n = 20 if s <= 4096 else 5
            o, l = fat.fwd(q, k, v, causal)
            tf = timeit(lambda: fat.fwd(q, k, v, causal), n)
            tb = timeit(lambda: fat.bwd(q, k, v, o, l, do, causal), max(2, n // 2))
            qt, kt, vt = [t.transpose(1, 2) for t in (q, k, v)]
            ts = timeit(lambda: F.scaled_dot_product_attention(qt, kt, vt, is_causal=causal), n)
            qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (qt, kt, vt))
            og = F.scaled_dot_product_attention(qg, kg, vg, is_causal=causal)
            dot = do.transpose(1, 2)
            tsb = timeit(lambda: torch.autograd.grad(og, (qg, kg, vg), dot, retain_graph=True), max(2, n // 2))
            fl = 4 * batch_size * num_heads * s * s * d * (0.5 if causal else 1.0)
            rf = rb = float("nan")
            if ref is not None and dt == torch.float16:
                ro, rl = ref.fwd(q, k, v, causal)
                rf = timeit(lambda: ref.fwd(q, k, v, causal), max(2, n // 4))
                rb = timeit(lambda: ref.bwd(q, k, v, ro, rl, do, causal), 2)
            print(f"| {s} | {d} | {causal} | {tf:.3f} | {fl/tf/1e9:.0f} | {tb:.3f} | {2.5*fl/tb/1e9:.0f} | {ts:.3f} | {fl/ts/1e9:.0f} | {tsb:.3f} | {2.5*fl/tsb/1e9:.0f} | {rf:.3f} | {rb:.3f} |", flush=True)
            del qg, kg, vg, og
