"""clock64 timeline of CTA (0,0,0) of the backward's main kernel (instrumented build, see scripts/trace_fwd.py):

    FA_B200_LIB=ab/trace/libfa_b200.so python scripts/trace_bwd.py [b s [head_dim]]

role 0 = elementwise thread 0, role 1 = MMA warp; events as stamped by FA_BTRACE in flash_bwd_tc_sm100.cu."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import cabi

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
s = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
torch.manual_seed(0)
q, k, v, do = (torch.randn(b, s, 32, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
o, lse = cabi.fwd(q, k, v, False)
for _ in range(2):
    cabi.bwd(q, k, v, o, lse, do, False)
torch.cuda.synchronize()
lib = cabi.load()
words = 3 * 64 * 8
buf = (ctypes.c_longlong * words)()
lib.fa_b200_trace_read.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
assert lib.fa_b200_trace_read(buf, words) == 0, "not a trace build"
t0 = min(x for x in buf if x > 0)
print(f"BTRACE bwd b{b} s{s} d{d}: role step : events (cycles since the first stamp)")
for r in range(2):
    for j in range(0, 12):
        ev = [buf[(r * 64 + j) * 8 + e] for e in range(8)]
        print(f"BTRACE {r} {j:2d} :" + "".join(f" {(x - t0) if x else -1:8d}" for x in ev))
