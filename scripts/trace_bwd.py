import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200"))
import torch, flash_attn_turing as fat
torch.manual_seed(0)
q = torch.randn(4, 4096, 32, 128, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q); do = torch.randn_like(q)
o, l = fat.fwd(q, k, v, False)
for _ in range(2): fat.bwd(q, k, v, o, l, do, False)
torch.cuda.synchronize()
os.environ["FA_B200_TRACE"] = "1"
fat.bwd(q, k, v, o, l, do, False)
torch.cuda.synchronize()
