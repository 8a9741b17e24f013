#!/bin/bash
# 8-GPU run of the driver's command: bench.py --gpus 8 (C2 weak headline + configs.C5shard = config 5 across the box)
L=gpurun_out/r02_n8.log
mkdir -p gpurun_out; : > $L
nvidia-smi --query-gpu=index,name --format=csv >> $L
nvidia-smi topo -m >> $L 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -c 600 gpurun_out/r02_bench_n8.err >> $L
python - >> $L <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/r02_bench_n8.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('N=8 value', round(j['value'],1), 'ms', round(j['ms_per_step'],3), j['config']['workload'], 'frac', round(j['roofline']['frac'],3), j['roofline']['peak_kind'], j['clocks'])
    for k,v in j.get("configs",{}).items(): print(k, {x: v.get(x) for x in ('value','ms_per_step','frac_of_burst_peak','frac_of_sustained_peak','clocks')})
    print('shard_io', j.get('shard_io'))
    print('e2e', {x: j['e2e'][x] for x in ('value','ms_per_step','copy_floor_ms','frac_of_copy_floor','numa_bound')})
except Exception as e: print('bench parse failed', e)
PY
tail -6 $L | cut -c1-600
