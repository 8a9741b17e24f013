#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee -a gpurun_out/run12.log) 2>&1
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
FA_B200_BWD=fused LD_LIBRARY_PATH=flash-attention-turing_b200/build/trace timeout 100 python scripts/trace_bwd.py > gpurun_out/trace_bwd_fused2.log 2>&1; grep "fused" gpurun_out/trace_bwd_fused2.log
