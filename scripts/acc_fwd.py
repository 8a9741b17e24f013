"""forward error of the selected kernel variant (FA_B200_FWD / FA_B200_EMU) against an fp32 reference on the GPU"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from gpu_diag import run_case
print("ACC variant", os.environ.get("FA_B200_FWD", "default"), "EMU", os.environ.get("FA_B200_EMU", "-"), flush=True)
for dt in (torch.bfloat16, torch.float16):
    run_case(1, 4, 4, 2048, 2048, 128, False, dt, do_bwd=False)
    run_case(1, 4, 2, 2048, 2048, 128, True, dt, do_bwd=False)
