"""Bit-reproducibility of the forward (and dK / dV of the backward) under a sustained loop: the first launch's outputs against those of
launches i = 1 .. N (back to back, power-capped clocks).

    python scripts/diag_repeat.py [--seconds S] b,s,h,d,causal ..."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import cabi

args = sys.argv[1:]
seconds = 1.0
if args and args[0] == "--seconds":
    seconds = float(args[1]); args = args[2:]
lib = cabi.load()
stream = torch.cuda.current_stream()
sptr = ctypes.c_void_p(stream.cuda_stream)
for spec in args:
    b, s, h, d, causal = (int(x) for x in spec.split(","))
    torch.manual_seed(0)
    q, k, v = (torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    o0, lse0 = cabi.fwd(q, k, v, bool(causal))
    o = torch.empty_like(q)
    lse = torch.empty_like(lse0)
    params = cabi.make_fwd_params(q, k, v, o, lse, bool(causal))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    lib.fa_b200_fwd(ctypes.byref(params), sptr)
    e1.record(stream)
    torch.cuda.synchronize()
    n_total = max(20, int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
    bad_o = bad_l = checks = 0
    done = 0
    while done < n_total:
        chunk = min(200, n_total - done)
        for _ in range(chunk):
            rc = lib.fa_b200_fwd(ctypes.byref(params), sptr)
            assert rc == 0
        done += chunk
        torch.cuda.synchronize()
        checks += 1
        if not torch.equal(o, o0):
            bad_o += 1
            if bad_o <= 2:
                diff = (o.float() - o0.float()).abs()
                idx = torch.nonzero(diff.amax(dim=-1) > 0)
                print(f"REPEAT   O differs after {done} launches: {idx.shape[0]} rows, max {diff.max().item():.3e}, first (b, row, head) {idx[0].tolist()} last {idx[-1].tolist()}", flush=True)
        if not torch.equal(lse, lse0):
            bad_l += 1
    print(f"REPEAT b{b} s{s} h{h} d{d} causal={causal}: {n_total} launches, {checks} checks, O mismatches {bad_o}, LSE mismatches {bad_l}", flush=True)
