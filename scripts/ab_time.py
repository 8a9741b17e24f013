"""A/B timing of one build of libfa_b200.so through the C ABI (device-resident tensors, CUDA events on the launching stream).

    FA_B200_LIB=<path to libfa_b200.so> python scripts/ab_time.py [--bwd] [--sustain SECONDS] [--dtype bf16|fp16] SHAPE ...

SHAPE = a BASELINE name (C2 C3 C4 C2c S1k S2k D64a D64c ...) or b,s,h,d,causal (e.g. 4,4096,32,128,0).
For every shape: a burst (n launches back to back after 3 warm-ups), optionally a sustained loop of SECONDS, and the max
abs error of batch 0 / last batch against torch's fused SDPA (a sanity check, not the parity suite).
"""
import argparse
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F

import cabi

SHAPES = {"C2": (4, 4096, 32, 128, 0), "C3": (4, 8192, 32, 128, 1), "C4": (4, 16384, 32, 128, 0), "C2c": (4, 4096, 32, 128, 1),
          "S1k": (16, 1024, 32, 128, 0), "S2k": (8, 2048, 32, 128, 0), "S512": (32, 512, 32, 128, 0),
          "D64a": (4, 4096, 32, 64, 0), "D64c": (4, 8192, 32, 64, 1), "D64l": (4, 16384, 32, 64, 0),
          "B4h16": (4, 4096, 16, 128, 0), "B4h16d64": (4, 4096, 16, 64, 0)}


def timeit(fn, n, stream):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shapes", nargs="*", default=["C2", "C3", "C4"])
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--sustain", type=float, default=0.0)
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--sdpa", action="store_true", help="also time torch's fused SDPA on the same tensors")
    args = ap.parse_args()
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype]
    lib = cabi.load()
    tag = os.environ.get("FA_TAG") or os.path.relpath(cabi.LIB_PATH, ROOT)
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)
    for nm in args.shapes:
        b, s, h, d, causal = SHAPES[nm] if nm in SHAPES else tuple(int(x) for x in nm.split(","))
        torch.manual_seed(0)
        q, k, v = (torch.randn(b, s, h, d, device="cuda", dtype=dt) for _ in range(3))
        o = torch.empty_like(q)
        lse = torch.empty(b, h, s, device="cuda", dtype=torch.float32)
        params = cabi.make_fwd_params(q, k, v, o, lse, bool(causal))

        def fwd():
            rc = lib.fa_b200_fwd(ctypes.byref(params), sptr)
            if rc != 0:
                raise RuntimeError(lib.fa_b200_last_error().decode())

        fl = 4.0 * b * h * s * s * d * (0.5 if causal else 1.0)
        n = args.iters or (20 if s <= 8192 else 8)
        ms = timeit(fwd, n, stream)
        errs = []
        for bi in (0, b - 1):
            r = F.scaled_dot_product_attention(q[bi:bi + 1].transpose(1, 2), k[bi:bi + 1].transpose(1, 2), v[bi:bi + 1].transpose(1, 2),
                                               is_causal=bool(causal)).transpose(1, 2)
            errs.append((o[bi:bi + 1].float() - r.float()).abs().max().item())
        line = (f"AB {tag} {nm} {args.dtype} b{b} s{s} h{h} d{d} c{causal} fwd burst n={n}: {ms:.4f} ms {fl / ms / 1e9:.1f} TF/s "
                f"| err vs SDPA {errs[0]:.2e} {errs[1]:.2e} finite={bool(torch.isfinite(o.float()).all())}")
        if args.sustain > 0:
            n2 = max(n, int(args.sustain * 1e3 / ms))
            ms2 = timeit(fwd, n2, stream)
            line += f" | sustained n={n2}: {ms2:.4f} ms {fl / ms2 / 1e9:.1f} TF/s"
        print(line, flush=True)
        if args.sdpa:
            qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))
            ms3 = timeit(lambda: F.scaled_dot_product_attention(qt, kt, vt, is_causal=bool(causal)), n, stream)
            print(f"AB torch-SDPA {nm} fwd burst n={n}: {ms3:.4f} ms {fl / ms3 / 1e9:.1f} TF/s", flush=True)
        if args.bwd:
            do = torch.randn_like(q)
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            dsum = torch.empty_like(lse)
            bp = cabi.BwdParams()
            bp.fwd = params
            bp.dout, bp.dq, bp.dk, bp.dv, bp.dsum = do.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dsum.data_ptr()
            nbytes = int(lib.fa_b200_bwd_workspace_bytes(ctypes.byref(bp.fwd)))
            ws = torch.empty(max(nbytes, 1), device="cuda", dtype=torch.uint8)
            bp.workspace = ws.data_ptr() if nbytes > 0 else None

            def bwd():
                rc = lib.fa_b200_bwd(ctypes.byref(bp), sptr)
                if rc != 0:
                    raise RuntimeError(lib.fa_b200_last_error().decode())

            nb = args.iters or (10 if s <= 8192 else 4)
            msb = timeit(bwd, nb, stream)
            # sanity: dV of one (batch, head) against autograd through torch's SDPA
            qs, ks, vs = (t[:1, :, :1].detach().clone().requires_grad_(True) for t in (q, k, v))
            r = F.scaled_dot_product_attention(qs.transpose(1, 2), ks.transpose(1, 2), vs.transpose(1, 2), is_causal=bool(causal)).transpose(1, 2)
            r.backward(do[:1, :, :1])
            e = [(a[:1, :, :1].float() - g_.float()).abs().max().item() for a, g_ in ((dq, qs.grad), (dk, ks.grad), (dv, vs.grad))]
            line = (f"AB {tag} {nm} {args.dtype} bwd burst n={nb}: {msb:.4f} ms {2.5 * fl / msb / 1e9:.1f} TF/s | err dq/dk/dv vs SDPA autograd "
                    f"{e[0]:.2e} {e[1]:.2e} {e[2]:.2e}")
            if max(e) > 0.1:   # who is wrong?  the same slice against the fp32 reference of the test suite
                from gpu_ref import attention_ref
                ref = attention_ref(q[:1, :, :1], k[:1, :, :1], v[:1, :, :1], bool(causal), do[:1, :, :1])
                e32 = [(a[:1, :, :1].float() - g_.float()).abs().max().item() for a, g_ in zip((dq, dk, dv), ref[2:])]
                s32 = [(g_.float() - r_.float()).abs().max().item() for g_, r_ in zip((qs.grad, ks.grad, vs.grad), ref[2:])]
                line += f" | vs fp32 reference: ours {e32[0]:.2e} {e32[1]:.2e} {e32[2]:.2e}, SDPA autograd {s32[0]:.2e} {s32[1]:.2e} {s32[2]:.2e}"
            print(line, flush=True)
            if args.sustain > 0:
                n2 = max(nb, int(args.sustain * 1e3 / msb))
                msb2 = timeit(bwd, n2, stream)
                print(f"AB {tag} {nm} bwd sustained n={n2}: {msb2:.4f} ms {2.5 * fl / msb2 / 1e9:.1f} TF/s", flush=True)
            del do, dq, dk, dv, ws
        del q, k, v, o, lse
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
