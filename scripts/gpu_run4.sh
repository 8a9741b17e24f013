#!/bin/bash
# short GPU session: p4 straight-line variant vs p4 base vs default; burst and sustained
set -u
mkdir -p gpurun_out
L=gpurun_out/run4.log
exec > >(tee -a $L) 2>&1
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
FA_B200_FWD=p4 timeout 60 python scripts/time_fwd.py S1k C2c C2 || { echo "P4 QUICK FAILED"; exit 1; }
echo "== default"
timeout 100 python scripts/time_fwd.py C2 C3 C4
echo "== p4 v_base (chunked, per-chunk mask branches)"
LD_LIBRARY_PATH=flash-attention-turing_b200/build/v_base FA_B200_FWD=p4 FA_B200_EMU=0 timeout 100 python scripts/time_fwd.py C2 C3
echo "== p4 straight-line"
FA_B200_FWD=p4 FA_B200_EMU=0 timeout 100 python scripts/time_fwd.py C2 C3 C4
FA_B200_FWD=p4 FA_B200_EMU=1 timeout 100 python scripts/time_fwd.py C2 C3 C4
echo "== sustained 200"
FA_ITERS=200 timeout 100 python scripts/time_fwd.py C2 C3
FA_ITERS=200 FA_B200_FWD=p4 FA_B200_EMU=0 timeout 100 python scripts/time_fwd.py C2 C3
FA_ITERS=200 FA_B200_FWD=p4 FA_B200_EMU=1 timeout 100 python scripts/time_fwd.py C2 C3
LD_LIBRARY_PATH=flash-attention-turing_b200/build/trace FA_B200_FWD=p4 timeout 100 python scripts/trace_fwd.py > gpurun_out/trace_p4c.log 2>&1; tail -45 gpurun_out/trace_p4c.log | head -32
echo "== done"
