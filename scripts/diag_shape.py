"""Per-head error report of one dense shape against the fp32 reference (tests/gpu_ref.py), batch by batch:

    python scripts/diag_shape.py b,s,h,d,causal[,hk] ...

Prints, for O / LSE / dQ / dK / dV, the largest absolute error and the (batch, head) it occurs at."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import cabi
from gpu_ref import attention_ref

for spec in sys.argv[1:]:
    v_ = [int(x) for x in spec.split(",")]
    b, s, h, d, causal = v_[:5]
    hk = v_[5] if len(v_) > 5 else h
    torch.manual_seed(0)
    q = torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(b, s, hk, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(b, s, hk, d, device="cuda", dtype=torch.bfloat16)
    do = torch.randn_like(q)
    o, lse = cabi.fwd(q, k, v, bool(causal))
    g = cabi.bwd(q, k, v, o, lse, do, bool(causal))
    worst = {n: (0.0, None) for n in ("o", "lse", "dq", "dk", "dv")}
    for bi in range(b):
        ref = attention_ref(q[bi:bi + 1], k[bi:bi + 1], v[bi:bi + 1], bool(causal), do[bi:bi + 1])
        ours = (o[bi:bi + 1], lse[bi:bi + 1], g[0][bi:bi + 1], g[1][bi:bi + 1], g[2][bi:bi + 1])
        for n, a, r in zip(worst, ours, ref):
            e = (a.float() - r.float()).abs()
            hd = 1 if n == "lse" else 2                       # head axis: lse [b,h,s], others [b,s,h,d]
            per_head = e.amax(dim=[i for i in range(e.dim()) if i != hd])
            m, hi = per_head.max(dim=0)
            if m.item() > worst[n][0]:
                worst[n] = (m.item(), (bi, int(hi)))
        del ref
    print(f"SHAPE b{b} s{s} h{h}/{hk} d{d} causal={causal}: " + "  ".join(f"{n} {e:.2e}@{w}" for n, (e, w) in worst.items()), flush=True)
