"""Random-shape stress of the forward (and backward) through the C ABI against the fp32 reference of tests/gpu_ref.py.

    python scripts/fuzz_shapes.py [N] [seed]

Shapes are drawn to hit what the fixed grid does not: many small items per CTA, one-tile items (sq <= 128), query tiles with
no visible key (causal, sq > sk), unequal block counts of the two query tiles, several waves of items, GQA, head_dim 64,
fp16 / bf16, score ranges that trigger the speculative path's retry.  A hang shows up as the caller's timeout."""
import math
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import cabi
from gpu_ref import attention_ref, assert_close

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = random.Random(seed)
torch.manual_seed(seed)
t0 = time.time()
bad = 0
verbose = os.environ.get("FUZZ_VERBOSE") == "1"     # print every shape BEFORE it runs (locating a hang)
for it in range(n):
    d = rng.choice([64, 128])
    dtype = rng.choice([torch.bfloat16, torch.float16])
    causal = rng.random() < 0.5
    hk = rng.choice([1, 2, 3, 4])
    h = hk * rng.choice([1, 1, 2, 4])
    kind = rng.random()
    if kind < 0.3:      # many tiny items
        b, sq, sk = rng.randint(8, 40), rng.randint(1, 300), rng.randint(1, 600)
    elif kind < 0.6:    # a few waves of medium items
        b, sq, sk = rng.randint(1, 6), rng.randint(200, 2500), rng.randint(1, 2500)
    elif kind < 0.8:    # sq > sk causal: leading query tiles see no key
        b, sq, sk = rng.randint(1, 4), rng.randint(300, 1500), rng.randint(1, 400)
    else:               # long keys, few rows
        b, sq, sk = rng.randint(1, 3), rng.randint(1, 260), rng.randint(2000, 9000)
    scale = rng.choice([1.0, 1.0, 4.0])    # 4.0: scores ~ N(0, 16 d) / sqrt(d): large jumps between key tiles
    q = (torch.randn(b, sq, h, d, device="cuda") * scale).to(dtype)
    k = (torch.randn(b, sk, hk, d, device="cuda") * scale).to(dtype)
    v = torch.randn(b, sk, hk, d, device="cuda").to(dtype)
    tag = f"#{it} b{b} sq{sq} sk{sk} h{h}/{hk} d{d} causal={causal} {dtype} scale={scale}"
    if verbose:
        print("FUZZ run", tag, flush=True)
    try:
        o, lse = cabi.fwd(q, k, v, causal)
        do_bwd = rng.random() < 0.4 and scale == 1.0
        dout = torch.randn_like(q) if do_bwd else None
        ref = attention_ref(q, k, v, causal, dout)
        assert_close(o, ref[0], dtype, "o")
        assert (lse - ref[1]).abs().max().item() < 2e-2 * max(1.0, ref[1].abs().max().item()), "lse"
        if do_bwd:
            dq, dk, dv = cabi.bwd(q, k, v, o, lse, dout, causal)
            assert_close(dq, ref[2], dtype, "dq")
            assert_close(dk, ref[3], dtype, "dk")
            assert_close(dv, ref[4], dtype, "dv")
        if verbose and hasattr(cabi.load(), "fa_b200_hang_read"):      # hang-guard build (-DFA_HANG_GUARD)
            import ctypes
            info = (ctypes.c_uint * 8)()
            torch.cuda.synchronize()
            if cabi.load().fa_b200_hang_read(info):
                print("FUZZ HANG (backward TU)", tag, "barrier", (info[0] - info[6]) // 8, "parity", info[1], "thread", info[2],
                      "block", info[3], info[4], info[5], flush=True)
                sys.exit(3)
    except AssertionError as e:
        bad += 1
        print("FUZZ FAIL", tag, str(e)[:300], flush=True)
    if it % 25 == 0:
        print(f"FUZZ progress {it}/{n} {tag} ({time.time() - t0:.0f}s)", flush=True)
print(f"FUZZ done: {n} shapes, {bad} failures, {time.time() - t0:.0f}s")
sys.exit(1 if bad else 0)
