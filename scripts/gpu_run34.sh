#!/bin/bash
# round 2, GPU call 34: validation of the shipping build (retry flag, fused head_dim-64 backward): GPU suite, smoke, bench line,
# head_dim-64 timings, ncu of the fused head_dim-64 backward, reference arm
O=gpurun_out/ckpt34; mkdir -p $O
L=$O/ckpt.log; : > $L
t0=$(date +%s); stamp() { echo "== $* (t+$(( $(date +%s) - t0 )) s)" >> $L; }
stamp "pytest -m gpu"
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $L
stamp smoke
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
stamp "bench (default flags)"
timeout 420 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/bench_n1.err >> $L; cut -c1-600 $O/bench_n1.json >> $L
stamp "head_dim 64 timings (default = fused backward) and det"
timeout 120 python scripts/ab_time.py --bwd --sustain 0.5 D64a D64c B4h16d64 4,16384,16,64,1 C2 >> $L 2>&1
FA_B200_BWD_D64=det FA_TAG=det timeout 100 python scripts/ab_time.py --bwd D64a B4h16d64 >> $L 2>&1
stamp "ncu --set full: backward head_dim 64"
timeout 200 ncu --set full --clock-control none -k regex:flash_bwd -s 12 -c 3 -o $O/bwd_d64 python scripts/ab_time.py --bwd --iters 2 D64a >> $L 2>&1
timeout 60 ncu -i $O/bwd_d64.ncu-rep --page raw --csv > $O/bwd_d64_raw.csv 2>> $L
stamp "reference arm"
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2>> $L
stamp done
grep -v "^==PROF\|^==WARN" $L | cut -c1-260 | tail -60
