#!/bin/bash
# round 2, GPU call 6: store warps + fast work-list decode; bench.py default run
L=gpurun_out/r02_run6.log
mkdir -p gpurun_out; : > $L
echo "== parity subset" >> $L
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_api_gpu.py -m gpu -x -q 2>&1 | tail -4 >> $L
echo "== A/B" >> $L
timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a S1k S512 >> $L 2>&1
FA_B200_EMU=3 FA_TAG="EMU=3" timeout 300 python scripts/ab_time.py C2 C3 D64a >> $L 2>&1
FA_B200_EMU=0 FA_TAG="EMU=0" timeout 300 python scripts/ab_time.py C2 C3 D64a S1k >> $L 2>&1
echo "== trace" >> $L
FA_B200_LIB=ab/trace/libfa_b200.so FA_TRACE_STEPS=28,36 timeout 120 python scripts/trace_fwd.py 4 4096 >> $L 2>&1
echo "== bench default" >> $L
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.err >> $L
python - >> $L <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
    print('bench value', round(j['value'],1), 'ms', round(j['ms_per_step'],4), 'frac', round(j['roofline']['frac'],3), j['clocks'])
    print('sustained', {k: j['sustained'][k] for k in ('value','ms_per_step','steps','frac_of_sustained_peak','clocks')})
    for k,v in j['configs'].items(): print(k, {x: v.get(x) for x in ('value','ms_per_step','frac_of_burst_peak','frac_of_sustained_peak','clocks','error')})
    print('e2e', j['e2e'])
    print('cpu', j['cpu_baseline'])
except Exception as e: print('bench parse failed', e)
PY
tail -3 $L
