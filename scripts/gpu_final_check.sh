#!/bin/bash
# validation of the committed defaults on one B200: GPU tests, smoke, the driver's bench line (all legs), the reference arm
#   gpurun --timeout 700 -- 'bash scripts/gpu_final_check.sh'     (~3 min of box time)
set -u
O=gpurun_out/final; mkdir -p $O
exec > >(tee -a $O/final.log) 2>&1
t0=$(date +%s); stamp() { echo "== $* (t+$(( $(date +%s) - t0 )) s)"; }
stamp "pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
stamp smoke
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"
stamp "bench (default flags)"
timeout 500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.err
python - <<PY
import json
j = json.loads(open("$O/bench_n1.json").read().strip().splitlines()[-1])
print("value", round(j["value"], 1), "frac", round(j["roofline"]["frac"], 3), "sustained", round(j["sustained"]["value"], 1),
      {k: round(v["value"], 1) for k, v in j["configs"].items()}, "e2e", round(j["e2e"]["value"], 1), j["clocks"])
PY
stamp "reference arm"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json; cut -c1-300 $O/bench_ref.json
stamp done
