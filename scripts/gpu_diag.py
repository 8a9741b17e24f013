"""Bring-up diagnostics run on the GPU box: prints error metrics for a sweep of shapes (fwd + bwd) against an fp32
torch reference evaluated on the GPU, plus environment facts.  Not a test; tests/ holds the asserted versions."""
import os, sys, time, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200"))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import flash_attn_turing as fat

dev = "cuda"
print("torch", torch.__version__, "dev", torch.cuda.get_device_name(0), "cpus", os.cpu_count(),
      "ref_dir_exists", os.path.exists("/root/reference"), flush=True)


def ref_attn(q, k, v, causal, dout=None):
    """fp32 reference on GPU: explicit softmax with bottom-right causal mask; rows with no key -> 0"""
    q32, k32, v32 = [t.float().permute(0, 2, 1, 3).detach().requires_grad_(True) for t in (q, k, v)]
    b, h, sq, d = q32.shape
    hk, sk = k32.shape[1], k32.shape[2]
    kk = k32.repeat_interleave(h // hk, dim=1)
    vv = v32.repeat_interleave(h // hk, dim=1)
    s = torch.matmul(q32, kk.transpose(-1, -2)) / math.sqrt(d)
    if causal:
        mask = torch.tril(torch.ones(sq, sk, dtype=torch.bool, device=q.device), diagonal=sk - sq)
        s = s.masked_fill(~mask, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.softmax(s, dim=-1)
    p = torch.where(p.isnan(), torch.zeros_like(p), p)
    lse = torch.where(torch.isinf(lse), torch.zeros_like(lse), lse)
    o = torch.matmul(p, vv)
    res = [o.permute(0, 2, 1, 3), lse]
    if dout is not None:
        g = torch.autograd.grad(o, (q32, k32, v32), dout.float().permute(0, 2, 1, 3))
        res += [x.permute(0, 2, 1, 3) for x in g]
    return res


def err(x, r):
    d = (x.float() - r.float()).abs()
    return f"max {d.max().item():.2e} mean {d.mean().item():.2e}"


def run_case(b, h, hk, sq, sk, d, causal, dtype, do_bwd=True):
    torch.manual_seed(1)
    q = torch.randn(b, sq, h, d, device=dev, dtype=dtype)
    k = torch.randn(b, sk, hk, d, device=dev, dtype=dtype)
    v = torch.randn(b, sk, hk, d, device=dev, dtype=dtype)
    do = torch.randn(b, sq, h, d, device=dev, dtype=dtype)
    tag = f"b{b} h{h}/{hk} {sq}x{sk} d{d} c{int(causal)} {str(dtype)[6:]}"
    try:
        o, l = fat.fwd(q, k, v, causal)
        torch.cuda.synchronize()
        r = ref_attn(q, k, v, causal, do if do_bwd else None)
        msg = f"{tag:44s} O {err(o, r[0])} | LSE {err(l, r[1])}"
        bad = (not torch.isfinite(o.float()).all().item())
        if do_bwd:
            dq, dk, dv = fat.bwd(q, k, v, o, l, do, causal)
            torch.cuda.synchronize()
            msg += f" | dQ {err(dq, r[2])} dK {err(dk, r[3])} dV {err(dv, r[4])}"
        print(msg + (" NONFINITE" if bad else ""), flush=True)
    except Exception as e:  # noqa
        print(f"{tag:44s} EXC {type(e).__name__}: {str(e)[:300]}", flush=True)
        raise


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    bf, hf = torch.bfloat16, torch.float16
    # smallest aligned case first: if descriptors are wrong this is where it shows
    run_case(1, 1, 1, 128, 128, 128, False, bf)
    run_case(1, 1, 1, 128, 128, 128, False, hf)
    run_case(1, 1, 1, 128, 256, 128, False, bf)
    run_case(1, 1, 1, 256, 512, 128, False, bf)
    run_case(1, 1, 1, 128, 128, 64, False, bf)
    run_case(1, 2, 1, 256, 384, 64, False, hf)
    run_case(1, 1, 1, 128, 128, 128, True, bf)
    run_case(2, 4, 2, 512, 512, 128, True, bf)
    run_case(1, 2, 1, 100, 100, 128, False, hf)
    run_case(1, 2, 1, 1, 1, 128, False, hf)
    run_case(3, 6, 3, 63, 65, 128, True, hf)
    run_case(3, 6, 1, 65, 63, 64, True, hf)
    run_case(1, 2, 1, 129, 127, 128, True, hf)
    run_case(1, 2, 1, 1025, 1023, 128, True, hf)
    run_case(1, 2, 1, 1023, 1025, 64, False, hf)
    run_case(2, 4, 4, 1024, 1024, 128, False, bf)
    if which == "all":
        # larger, fwd only vs torch SDPA (fp32 reference would be too big): compare with bf16 SDPA
        for (b, s, causal) in [(1, 4096, False), (1, 4096, True)]:
            torch.manual_seed(2)
            q = torch.randn(b, s, 8, 128, device=dev, dtype=bf); k = torch.randn_like(q); v = torch.randn_like(q)
            o, l = fat.fwd(q, k, v, causal)
            r = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal).transpose(1, 2)
            print(f"b{b} s{s} h8 d128 c{int(causal)} vs torch SDPA bf16: O {err(o, r)}", flush=True)
        # timing: BASELINE config 2 / 3
        for (b, s, causal) in [(4, 4096, False), (4, 8192, True), (4, 16384, False)]:
            torch.manual_seed(0)
            q = torch.randn(b, s, 32, 128, device=dev, dtype=bf); k = torch.randn_like(q); v = torch.randn_like(q)
            for _ in range(3): fat.fwd(q, k, v, causal)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for _ in range(n): fat.fwd(q, k, v, causal)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            fl = 4 * b * 32 * s * s * 128 * (0.5 if causal else 1.0)
            print(f"TIMING b{b} s{s} causal={causal}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
            qt, kt, vt = [t.transpose(1, 2) for t in (q, k, v)]
            for _ in range(3): F.scaled_dot_product_attention(qt, kt, vt, is_causal=causal)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n): F.scaled_dot_product_attention(qt, kt, vt, is_causal=causal)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print(f"TIMING torch SDPA same shape: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    print("DIAG DONE", flush=True)
