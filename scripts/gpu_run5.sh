#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/run5.log
exec > >(tee -a $L) 2>&1
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
FA_B200_FWD=p4 FA_B200_EMU=3 timeout 60 python scripts/time_fwd.py S1k C2c C2 || { echo "P4 QUICK FAILED"; exit 1; }
echo "== accuracy vs fp32"
timeout 100 python scripts/acc_fwd.py 2>&1 | grep -v "^torch"
FA_B200_FWD=p4 FA_B200_EMU=0 timeout 100 python scripts/acc_fwd.py 2>&1 | grep -v "^torch"
FA_B200_FWD=p4 FA_B200_EMU=1 timeout 100 python scripts/acc_fwd.py 2>&1 | grep -v "^torch"
FA_B200_FWD=p4 FA_B200_EMU=3 timeout 100 python scripts/acc_fwd.py 2>&1 | grep -v "^torch"
echo "== default"
timeout 100 python scripts/time_fwd.py C2 C3 C4
echo "== p4 v_base EMU1"
LD_LIBRARY_PATH=flash-attention-turing_b200/build/v_base FA_B200_FWD=p4 FA_B200_EMU=1 timeout 100 python scripts/time_fwd.py C2 C3 C4
echo "== p4 straight-line EMU 1 / 3 / 2"
FA_B200_FWD=p4 FA_B200_EMU=1 timeout 100 python scripts/time_fwd.py C2 C3 C4
FA_B200_FWD=p4 FA_B200_EMU=3 timeout 100 python scripts/time_fwd.py C2 C3 C4
FA_B200_FWD=p4 FA_B200_EMU=2 timeout 100 python scripts/time_fwd.py C2 C3
echo "== pytest p4 EMU1"
FA_B200_FWD=p4 FA_B200_EMU=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== done"
