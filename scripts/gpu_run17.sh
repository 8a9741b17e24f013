#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee -a gpurun_out/run17.log) 2>&1
B=flash-attention-turing_b200/build
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
echo "== red16"; timeout 120 python scripts/time_bwd.py S2k C2 C3 || { echo "QUICK FAILED"; exit 1; }
echo "== red32"; LD_LIBRARY_PATH=$B/v_red32 timeout 120 python scripts/time_bwd.py C2 C3
echo "== red0";  LD_LIBRARY_PATH=$B/v_red0 timeout 120 python scripts/time_bwd.py C2
echo "== red16 C4"; timeout 200 python scripts/time_bwd.py C4
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
LD_LIBRARY_PATH=$B/trace timeout 100 python scripts/trace_bwd.py 2>&1 | grep "fused 1"
