#!/bin/bash
# 2-GPU validation: NCCL shard I/O test, second-device test, bench.py --gpus 2 (C2 weak headline, configs.C5shard, shard_io, e2e)
L=gpurun_out/r02_multi.log
mkdir -p gpurun_out; : > $L
nvidia-smi --query-gpu=index,name --format=csv >> $L
echo "== pytest multi-GPU tests" >> $L
timeout 900 python -m pytest tests/test_shard_gpu.py tests/test_api_gpu.py -m gpu -x -q -k "nccl or second_device or threads or graph" 2>&1 | tail -6 >> $L
echo "== bench --gpus 2" >> $L
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 800 gpurun_out/r02_bench_n2.err >> $L
python - >> $L <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/r02_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('N=2 value', round(j['value'],1), 'ms', round(j['ms_per_step'],3), j['config']['workload'], 'frac', round(j['roofline']['frac'],3), j['roofline']['peak_kind'], j['clocks'])
    for k,v in j.get("configs",{}).items(): print(k, {x: v.get(x) for x in ('value','ms_per_step','frac_of_burst_peak','clocks')})
    print('shard_io', j.get('shard_io'))
    print('e2e', {x: j['e2e'][x] for x in ('value','ms_per_step','copy_floor_ms','frac_of_copy_floor','numa_bound')})
except Exception as e: print('bench parse failed', e)
PY
echo "== reference arm" >> $L
timeout 300 python bench.py --impl reference --gpus 2 --steps 5 2>&1 | tail -1 | cut -c1-400 >> $L
tail -3 $L
