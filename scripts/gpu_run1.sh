#!/bin/bash
# GPU session of the 4-softmax-warpgroup forward (p4): correctness (fail fast), timing matrix, clock64 trace, tests, bench
set -u
mkdir -p gpurun_out
L=gpurun_out/run2.log
exec > >(tee -a $L) 2>&1
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,clocks.sm --format=csv
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
echo "== p4 quick correctness"
FA_B200_FWD=p4 timeout 60 python scripts/time_fwd.py S1k C2c C2
rc=$?
if [ $rc -ne 0 ]; then
  echo "P4 QUICK FAILED rc=$rc -- falling back to default-kernel work only"
  P4OK=0
else
  P4OK=1
fi
echo "== timing matrix (burst)"
FA_TIME_SDPA=1 timeout 120 python scripts/time_fwd.py C2 C3
if [ $P4OK -eq 1 ]; then
  for emu in 0 1 2; do
    FA_B200_FWD=p4 FA_B200_EMU=$emu timeout 120 python scripts/time_fwd.py C2 C3 C4 || echo "P4 EMU $emu FAILED"
  done
  echo "== sustained (200 launches)"
  FA_ITERS=200 FA_B200_FWD=p4 timeout 120 python scripts/time_fwd.py C2
  FA_ITERS=200 FA_B200_FWD=p4 FA_B200_EMU=1 timeout 120 python scripts/time_fwd.py C2
  echo "== trace p4"
  LD_LIBRARY_PATH=flash-attention-turing_b200/build/trace FA_B200_FWD=p4 timeout 100 python scripts/trace_fwd.py > gpurun_out/trace_p4.log 2>&1; tail -45 gpurun_out/trace_p4.log
  echo "== pytest gpu with p4"
  FA_B200_FWD=p4 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
  FA_B200_FWD=p4 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_p4.json 2> gpurun_out/bench_p4.err; cat gpurun_out/bench_p4.json; tail -3 gpurun_out/bench_p4.err
fi
echo "== ubench"
timeout 60 ./scripts/ubench_softmax.bin
echo "== bench default"
timeout 400 python bench.py --bwd > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== done"
