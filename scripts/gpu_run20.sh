#!/bin/bash
# round 2, GPU call 20: checkpoint of the MMA2 + 1-in-8 polynomial defaults — full GPU suite, smoke, bench N=1, ncu of the forward
O=gpurun_out/ckpt20; mkdir -p $O
L=$O/ckpt.log; : > $L
echo "== pytest -m gpu" >> $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
echo "== smoke" >> $L
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "== bench (default flags)" >> $L
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.err >> $L
echo "== reference arm" >> $L
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2>> $L
echo "== launch list of the bench command" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_under_ncu.json 2>> $L
echo "== ncu --set full: forward C2 / C3 / d64" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_fwd -s 3 -c 1 -o $O/fwd_c2 python scripts/ab_time.py --iters 2 C2 >> $L 2>&1
timeout 600 ncu --set full --clock-control none -k regex:flash_fwd -s 3 -c 1 -o $O/fwd_c3 python scripts/ab_time.py --iters 2 C3 >> $L 2>&1
timeout 600 ncu --set full --clock-control none -k regex:flash_fwd -s 3 -c 1 -o $O/fwd_d64 python scripts/ab_time.py --iters 2 D64a >> $L 2>&1
grep -v "^==PROF\|^==WARN" $L | tail -30 | cut -c1-300
