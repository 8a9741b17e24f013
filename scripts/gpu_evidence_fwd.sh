#!/bin/bash
# (round-1 evidence script, kept for the history of profiles/r01*: it uses the bench flags of that time — `--bwd` is gone, the backward legs are configs.C4bwd / D64bwd of the default line now; see scripts/gpu_final_check.sh)
# evidence run: bench lines (C2 default + bwd, C3, C4fwd + bwd, 200-step sustained), ncu launch list of the bench command,
# one ncu --set full capture of the forward kernel
set -u
mkdir -p gpurun_out
L=gpurun_out/run8.log
exec > >(tee -a $L) 2>&1
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
timeout 60 python scripts/time_fwd.py C2 || { echo "QUICK FAILED"; exit 1; }
echo "== bench C2 default (+bwd)"
timeout 400 python bench.py --bwd > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json; tail -2 gpurun_out/bench_c2.err
echo "== bench reference arm"
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
echo "== bench C2 200 steps"
timeout 300 python bench.py --steps 200 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_200.json 2>/dev/null; cat gpurun_out/bench_c2_200.json
echo "== bench C3 (+bwd)"
timeout 300 python bench.py --config C3 --bwd --no-cpu-baseline --no-e2e > gpurun_out/bench_c3.json 2>/dev/null; cat gpurun_out/bench_c3.json
echo "== bench C4fwd (+bwd)"
timeout 400 python bench.py --config C4fwd --steps 10 --bwd --no-cpu-baseline --no-e2e > gpurun_out/bench_c4.json 2>/dev/null; cat gpurun_out/bench_c4.json
echo "== ncu launch list"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --bwd --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.json 2>/dev/null
grep -c flash gpurun_out/launches_bench.csv
echo "== ncu full fwd"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:flash_fwd -s 3 -c 1 -o gpurun_out/prof_fwd_p4 python scripts/time_fwd.py C2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
echo "== done"
