#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/run7.log
exec > >(tee -a $L) 2>&1
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
timeout 60 python scripts/time_fwd.py S1k C2c C2 || { echo "QUICK FAILED"; exit 1; }
echo "== pytest gpu (defaults: p4 sumvote EMU1)"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
echo "== timing"
FA_TIME_SDPA=1 timeout 100 python scripts/time_fwd.py C2 C3 C4
FA_B200_FWD=ws timeout 100 python scripts/time_fwd.py C2 C3 C4
FA_ITERS=200 timeout 100 python scripts/time_fwd.py C2 C3
FA_ITERS=200 FA_B200_FWD=ws timeout 100 python scripts/time_fwd.py C2 C3
echo "== trace"
LD_LIBRARY_PATH=flash-attention-turing_b200/build/trace timeout 100 python scripts/trace_fwd.py > gpurun_out/trace_p4_final.log 2>&1; tail -45 gpurun_out/trace_p4_final.log | head -31
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"
echo "== done"
