#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/run11.log
exec > >(tee -a $L) 2>&1
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
FA_B200_BWD=fused timeout 90 python scripts/time_bwd.py S2k C2 || { echo "FUSED QUICK FAILED rc=$?"; exit 1; }
timeout 120 python scripts/time_bwd.py C2
FA_B200_BWD=fused timeout 200 python scripts/time_bwd.py C2c C3 C4
FA_B200_BWD=fused timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
FA_B200_BWD=fused LD_LIBRARY_PATH=flash-attention-turing_b200/build/trace timeout 100 python scripts/trace_bwd.py > gpurun_out/trace_bwd_fused.log 2>&1; grep "fused" gpurun_out/trace_bwd_fused.log
echo "== done"
