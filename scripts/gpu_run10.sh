#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/run10.log
exec > >(tee -a $L) 2>&1
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
timeout 120 python scripts/time_bwd.py C2 || { echo "BWD QUICK FAILED"; exit 1; }
FA_B200_BWD_EMU=0 timeout 120 python scripts/time_bwd.py C2 C3
timeout 200 python scripts/time_bwd.py C3 C4
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
LD_LIBRARY_PATH=flash-attention-turing_b200/build/trace timeout 100 python scripts/trace_bwd.py > gpurun_out/trace_bwd_s3.log 2>&1; grep "dq" gpurun_out/trace_bwd_s3.log
echo "== done"
