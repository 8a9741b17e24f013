#!/bin/bash
# round 2, GPU call 11: deferred ring release variant; full GPU suite; bench N=1
L=gpurun_out/r02_run11.log
mkdir -p gpurun_out; : > $L
for lib in flash-attention-turing_b200/flash_attn_turing/libfa_b200.so ab/defer/libfa_b200.so; do
  echo "== A/B $lib" >> $L
  FA_B200_LIB=$lib timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a S1k >> $L 2>&1
done
echo "== pytest gpu (all)" >> $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $L
echo "== bench default" >> $L
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.err >> $L
python - >> $L <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
    print('bench value', round(j['value'],1), 'ms', round(j['ms_per_step'],4), 'frac', round(j['roofline']['frac'],3), j['clocks'])
    print('sustained', {k: j['sustained'][k] for k in ('value','ms_per_step','steps','frac_of_sustained_peak','clocks')})
    for k,v in j['configs'].items(): print(k, {x: v.get(x) for x in ('value','ms_per_step','frac_of_burst_peak','frac_of_sustained_peak','clocks','error')})
    print('e2e', {x: j['e2e'][x] for x in ('value','ms_per_step','copy_floor_ms','frac_of_copy_floor','gpu_launches_per_step')})
except Exception as e: print('bench parse failed', e)
PY
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
tail -3 $L
