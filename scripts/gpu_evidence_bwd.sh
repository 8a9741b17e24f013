#!/bin/bash
# (round-1 evidence script, kept for the history of profiles/r01*: it uses the bench flags of that time — `--bwd` is gone, the backward legs are configs.C4bwd / D64bwd of the default line now; see scripts/gpu_final_check.sh)
set -u
mkdir -p gpurun_out
exec > >(tee -a gpurun_out/run13.log) 2>&1
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
timeout 120 python scripts/time_bwd.py C2 || { echo "QUICK FAILED"; exit 1; }
echo "== pytest gpu (defaults: p4 fwd, fused bwd)"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
echo "== pytest gpu FA_B200_BWD=det"
FA_B200_BWD=det timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"
echo "== bench"
timeout 400 python bench.py --bwd --no-cpu-baseline > gpurun_out/bench_c2_fusedbwd.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/bench_c2_fusedbwd.json')); print('C2 fwd', j['value'], 'bwd', j['bwd'], 'e2e', j['e2e']['ms_per_step'])"
FA_B200_BWD=det timeout 400 python bench.py --bwd --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_detbwd.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/bench_c2_detbwd.json')); print('C2 det bwd', j['bwd']['ms_per_step'], j['bwd']['value'])"
timeout 400 python bench.py --config C3 --bwd --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_fusedbwd.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/bench_c3_fusedbwd.json')); print('C3 fwd', j['value'], 'bwd', j['bwd']['ms_per_step'], j['bwd']['value'])"
timeout 400 python bench.py --config C4fwd --steps 10 --bwd --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_fusedbwd.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/bench_c4_fusedbwd.json')); print('C4 fwd', j['value'], 'bwd', j['bwd']['ms_per_step'], j['bwd']['value'])"
echo "== ncu launch list (bench --bwd)"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bench2.csv python bench.py --steps 5 --warmup 3 --bwd --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -c flash gpurun_out/launches_bench2.csv
echo "== ncu full bwd fused"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:flash_bwd_dk_dv -s 2 -c 1 -o gpurun_out/prof_bwd_fused python scripts/time_bwd.py C2 > /dev/null 2>&1
ls -la gpurun_out/prof_bwd_fused.ncu-rep
echo "== done"
