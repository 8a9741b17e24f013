#!/bin/bash
# round 2, GPU call 41: the default bench line with the head_dim-64 legs (configs.D64fwd / D64bwd)
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
tail -c 300 gpurun_out/r02c_bench_n1.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02c_bench_n1.json").read().strip().splitlines()[-1])
print("value", round(j["value"], 1), "sustained", round(j["sustained"]["value"], 1), {k: (round(v["value"], 1), round(v["ms_per_step"], 3)) for k, v in j["configs"].items()})
print("e2e", j["e2e"]["value"], j["e2e"]["ms_per_step"], "cpu", j["cpu_baseline"]["value"], "clocks", j["clocks"])
PY
