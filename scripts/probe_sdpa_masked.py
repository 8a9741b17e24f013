import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200"))
import torch, torch.nn.functional as F
import flash_attn_turing as fat
torch.manual_seed(0)
sq, sk, h, hk, d = 128, 64, 2, 1, 64
q = torch.randn(1, sq, h, d, device="cuda", dtype=torch.float16); k = torch.randn(1, sk, hk, d, device="cuda", dtype=torch.float16); v = torch.randn_like(k)
mask = torch.tril(torch.ones(sq, sk, dtype=torch.bool, device="cuda"), diagonal=sk - sq)
ref = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=mask, enable_gqa=True).transpose(1, 2)
o, l = fat.fwd(q, k, v, True)
print("PROBE ref rows 0..3 (fully masked) abs max:", ref[0, :4].abs().amax(dim=(1, 2)).tolist(), "nan?", torch.isnan(ref).any().item())
print("PROBE ours rows 0..3 abs max:", o[0, :4].abs().amax(dim=(1, 2)).tolist())
diff = (o.float() - ref.float()).abs().amax(dim=(0, 2, 3))
print("PROBE per-row max diff: first masked rows", diff[:4].tolist(), " row 63,64,65:", diff[63:66].tolist(), " last:", diff[-2:].tolist())
print("PROBE mean of V:", v.float().mean(dim=1)[0, 0, :4].tolist(), " ref row0:", ref[0, 0, 0, :4].tolist())
# which backend
from torch.nn.attention import SDPBackend, sdpa_kernel
for be in (SDPBackend.MATH, SDPBackend.EFFICIENT_ATTENTION):
    try:
        with sdpa_kernel(be):
            r2 = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=mask, enable_gqa=True).transpose(1, 2)
        print("PROBE backend", be, "row0 absmax", r2[0, 0].abs().max().item(), "nan", torch.isnan(r2).any().item(), "maxdiff vs ours on rows>=64", (r2[0, 64:].float() - o[0, 64:].float()).abs().max().item())
    except Exception as e:
        print("PROBE backend", be, "failed", str(e)[:100])
