"""debug: per-row forward error for inputs whose scores grow along the keys (the lazy-rescale slow path)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, flash_attn_turing as fat
from gpu_ref import attention_ref
tag = f"FWD={os.environ.get('FA_B200_FWD','default')} EMU={os.environ.get('FA_B200_EMU','-')}"
dt = torch.bfloat16
torch.manual_seed(7)
b, sq, sk, h, hk, d = 2, 300, 1024, 4, 2, 128
q = torch.randn(b, sq, h, d, device="cuda", dtype=dt); k = torch.randn(b, sk, hk, d, device="cuda", dtype=dt)
v = torch.randn(b, sk, hk, d, device="cuda", dtype=dt)
grow = (1.0 + 6.0 * (torch.arange(sk, device="cuda") // 128)).to(dt)
k = (k * grow[None, :, None, None]).contiguous()
o, lse = fat.fwd(q, k, v, False)
ro, rl = attention_ref(q, k, v, False)
bi, qi, hi = 0, 290, 3
s = (q[bi, qi, hi].float() @ k[bi, :, hi // 2].float().T) / d ** 0.5
tile_max = s.view(8, 128).amax(dim=1)
print(f"DBG {tag} lse ours {lse[bi, hi, qi].item():.4f} ref {rl[bi, hi, qi].item():.4f} | tile maxima {[round(x, 1) for x in tile_max.tolist()]}")
print(f"DBG {tag} O[:4] ours {o[bi, qi, hi, :4].float().tolist()} ref {ro[bi, qi, hi, :4].tolist()} v833 {v[bi, 833, hi // 2, :4].float().tolist()} v953 {v[bi, 953, hi // 2, :4].float().tolist()}")
err = (o.float() - ro).abs().amax(dim=-1)
print(f"DBG {tag} bad rows {(err > 0.05).nonzero().tolist()}")
# same row alone, and with the big key moved into another column / tile
for shift in (0, 1, 64, 128):
    k2 = torch.roll(k, shifts=shift, dims=1) if shift else k
    v2 = torch.roll(v, shifts=shift, dims=1) if shift else v
    o2, l2 = fat.fwd(q[:1, 256:300].contiguous(), k2[:1], v2[:1], False)
    r2o, r2l = attention_ref(q[:1, 256:300].contiguous(), k2[:1], v2[:1], False)
    e2 = (o2.float() - r2o).abs().amax(dim=-1)
    print(f"DBG {tag} rows 256..299 alone, keys rolled by {shift}: max err {e2.max().item():.3e} lse err {(l2 - r2l).abs().max().item():.3e}")
