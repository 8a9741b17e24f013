"""clock64 timeline of CTA 0 of the forward kernel.  Needs the instrumented build:

    make -C flash-attention-turing_b200 trace          # -> ab/trace/libfa_b200.so (-DFA_TRACE)
    FA_B200_LIB=ab/trace/libfa_b200.so python scripts/trace_fwd.py [b s [d [causal]]]

Rows (one per key step): role 0 / 1 = softmax warp 0 of query tile 0 / 1 (column half 0, lane quadrant 0), role 2 / 3 =
softmax warp 7 of the tile (column half 1, quadrant 3), role 4 / 5 = the MMA warp's work for tile 0 / 1.
softmax events: 0 S_t full seen, 1 row max exchanged, 2..5 P quarter 0..3 handed over, 6 O_t full (item end), 7 epilogue done.
Rows 6 / 7: epilogue of item j (tile 0 / 1, warp 0): 0 start, 1 O full, 2 l exchanged, 3 staging tile taken, 4 O in registers,
5 staged, 6 handed over.  MMA events: 0..3 P quarter q available, 4 last P V issued, 5 next K available, 6 next S issued.  Cycles since the first stamp."""
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
causal = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
lo, hi = (int(x) for x in os.environ.get("FA_TRACE_STEPS", "24,40").split(","))
torch.manual_seed(0)
q, k, v = (torch.randn(b, s, 32, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
for _ in range(3):
    cabi.fwd(q, k, v, causal)
torch.cuda.synchronize()
cabi.fwd(q, k, v, causal)
lib = cabi.load()
words = 8 * 64 * 8
buf = (ctypes.c_longlong * words)()
lib.fa_b200_trace_read.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
assert lib.fa_b200_trace_read(buf, words) == 0, "not a trace build"
t0 = min(x for x in buf if x > 0)
print(f"TRACE fwd b{b} s{s} d{d} causal={causal}: role step : events (cycles since the first stamp)")
for r in range(8):
    for j in (range(lo, hi) if r < 6 else range(0, 4)):
        ev = [buf[(r * 64 + j) * 8 + e] for e in range(8)]
        print(f"TRACE {r} {j:2d} :" + "".join(f" {(x - t0) if x else -1:8d}" for x in ev))
