import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200"))
import torch, flash_attn_turing as fat
torch.manual_seed(0)
b, s = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 4096
q = torch.randn(b, s, 32, 128, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
for _ in range(3): fat.fwd(q, k, v, False)
torch.cuda.synchronize()
os.environ["FA_B200_TRACE"] = "1"
fat.fwd(q, k, v, False)
torch.cuda.synchronize()
