"""quick forward timing of the BASELINE shapes (device-resident, CUDA events); usage: time_fwd.py [C2 C3 C4]
env: FA_B200_FWD / FA_B200_EMU select the kernel variant, FA_ITERS overrides the timed launches per shape,
FA_TIME_SDPA=1 also times torch's fused SDPA on the same tensors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200"))
import torch, flash_attn_turing as fat
import torch.nn.functional as F
shapes = {"C2": (4, 4096, False), "C3": (4, 8192, True), "C4": (4, 16384, False), "C2c": (4, 4096, True), "S1k": (16, 1024, False), "S2k": (8, 2048, False)}
names = sys.argv[1:] or ["C2", "C3", "C4"]
tag = f"FWD={os.environ.get('FA_B200_FWD', 'default')} EMU={os.environ.get('FA_B200_EMU', '-')}"


def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for nm in names:
    b, s, causal = shapes[nm]
    torch.manual_seed(0)
    q = torch.randn(b, s, 32, 128, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
    o, l = fat.fwd(q, k, v, causal)
    n = int(os.environ.get("FA_ITERS", "20" if s <= 8192 else "8"))
    ms = timeit(lambda: fat.fwd(q, k, v, causal), n)
    fl = 4 * b * 32 * s * s * 128 * (0.5 if causal else 1.0)
    r = F.scaled_dot_product_attention(q[:1].transpose(1, 2), k[:1].transpose(1, 2), v[:1].transpose(1, 2), is_causal=causal).transpose(1, 2)
    err = (o[:1].float() - r.float()).abs()
    rl = F.scaled_dot_product_attention(q[-1:].transpose(1, 2), k[-1:].transpose(1, 2), v[-1:].transpose(1, 2), is_causal=causal).transpose(1, 2)
    errl = (o[-1:].float() - rl.float()).abs()
    print(f"TIMING {nm} {tag} n={n} b{b} s{s} causal={causal}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s | vs SDPA max {err.max().item():.2e} "
          f"mean {err.mean().item():.2e} last-batch max {errl.max().item():.2e} finite={bool(torch.isfinite(o.float()).all())}", flush=True)
    if os.environ.get("FA_TIME_SDPA") == "1":
        qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))
        ms2 = timeit(lambda: F.scaled_dot_product_attention(qt, kt, vt, is_causal=causal), n)
        print(f"TIMING {nm} torch-SDPA n={n}: {ms2:.3f} ms  {fl / ms2 / 1e9:.1f} TFLOP/s", flush=True)
