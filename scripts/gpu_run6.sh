#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/run6.log
exec > >(tee -a $L) 2>&1
VS=flash-attention-turing_b200/build/v_sum
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
LD_LIBRARY_PATH=$VS FA_B200_FWD=p4 FA_B200_EMU=1 timeout 60 python scripts/time_fwd.py S1k C2c C2 || { echo "P4 SUM QUICK FAILED"; exit 1; }
echo "== default"
timeout 100 python scripts/time_fwd.py C2 C3 C4
echo "== p4 maxvote EMU 1 / 4"
FA_B200_FWD=p4 FA_B200_EMU=1 timeout 100 python scripts/time_fwd.py C2 C3 C4
FA_B200_FWD=p4 FA_B200_EMU=4 timeout 100 python scripts/time_fwd.py C2 C3 C4
echo "== p4 sumvote EMU 0 / 1 / 4"
LD_LIBRARY_PATH=$VS FA_B200_FWD=p4 FA_B200_EMU=0 timeout 100 python scripts/time_fwd.py C2 C3
LD_LIBRARY_PATH=$VS FA_B200_FWD=p4 FA_B200_EMU=1 timeout 100 python scripts/time_fwd.py C2 C3 C4
LD_LIBRARY_PATH=$VS FA_B200_FWD=p4 FA_B200_EMU=4 timeout 100 python scripts/time_fwd.py C2 C3 C4
echo "== pytest p4 sumvote EMU1"
LD_LIBRARY_PATH=$VS FA_B200_FWD=p4 FA_B200_EMU=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
LD_LIBRARY_PATH=$VS FA_B200_FWD=p4 FA_B200_EMU=1 timeout 100 python scripts/acc_fwd.py 2>&1 | grep -v "^torch"
echo "== done"
