#!/bin/bash
# final validation of the committed defaults on one B200: GPU tests, smoke, the driver's bench line, the C5 shard config
set -u
mkdir -p gpurun_out
exec > >(tee -a gpurun_out/final.log) 2>&1
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')"
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"
echo "== bench (default flags)"
timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.json | cut -c1-1500
echo "== bench --bwd C4"
timeout 500 python bench.py --config C4fwd --steps 10 --bwd --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('C4 fwd', round(j['value'],1), 'bwd', round(j['bwd']['ms_per_step'],2), round(j['bwd']['value'],1), j['clocks'])"
echo "== C5 shard (b=32 s=16384), 2 steps"
timeout 500 python bench.py --config C5shard --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('C5shard fwd', round(j['value'],1), 'ms', round(j['ms_per_step'],2), j['clocks'])"
echo "== done"
