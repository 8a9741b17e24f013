#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee -a gpurun_out/run15.log) 2>&1
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
FA_B200_BWD_ACC=16 timeout 120 python scripts/time_bwd.py S2k C2 || { echo "ACC16 QUICK FAILED"; exit 1; }
timeout 120 python scripts/time_bwd.py C2
FA_B200_BWD_ACC=16 timeout 200 python scripts/time_bwd.py C3 C4
FA_B200_BWD_ACC=16 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
FA_B200_BWD_ACC=16 LD_LIBRARY_PATH=flash-attention-turing_b200/build/trace timeout 100 python scripts/trace_bwd.py 2>&1 | grep "fused 0"
