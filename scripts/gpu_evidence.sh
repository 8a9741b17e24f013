#!/bin/bash
# Evidence for profiles/ (one B200): ncu launch list of the bench command + full captures of every kernel, the
# benchmark.sh-grid sweep (fp16 + bf16, head_dim 64 / 128), the reference's full acceptance matrix for this module and for
# the reference's own kernels rebuilt for sm_100a.
O=gpurun_out/evidence2; mkdir -p $O
L=$O/evidence.log; : > $L
echo "== launch list of the bench command (shares, cold-cache)" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_under_ncu.json 2>> $L
echo "== ncu --set full: forward C2 / C3 / d64" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_fwd -s 3 -c 1 -o $O/fwd_c2 python scripts/ab_time.py --iters 2 C2 >> $L 2>&1
timeout 600 ncu --set full --clock-control none -k regex:flash_fwd -s 3 -c 1 -o $O/fwd_c3 python scripts/ab_time.py --iters 2 C3 >> $L 2>&1
timeout 600 ncu --set full --clock-control none -k regex:flash_fwd -s 3 -c 1 -o $O/fwd_d64 python scripts/ab_time.py --iters 2 D64a >> $L 2>&1
echo "== ncu --set full: backward C2 (dot, fused, convert) and d64 (dq, dk_dv)" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_bwd -s 6 -c 3 -o $O/bwd_c2 python scripts/ab_time.py --iters 2 --bwd C2 >> $L 2>&1
timeout 900 ncu --set full --clock-control none -k regex:flash_bwd -s 6 -c 3 -o $O/bwd_d64 python scripts/ab_time.py --iters 2 --bwd D64a >> $L 2>&1
echo "== sweep fp16" >> $L
timeout 900 python scripts/benchmark_sweep.py --dtype fp16 > $O/sweep_fp16.md 2>> $L
echo "== sweep bf16" >> $L
timeout 900 python scripts/benchmark_sweep.py --dtype bf16 > $O/sweep_bf16.md 2>> $L
echo "== reference's full acceptance matrix, math SDPA backend: ours, then the reference's own kernels" >> $L
timeout 1200 python scripts/run_reference_tests.py 1 ours math > $O/refmatrix_ours.log 2>&1
timeout 1200 python scripts/run_reference_tests.py 1 ref math > $O/refmatrix_ref.log 2>&1
grep REFTEST $O/refmatrix_ours.log $O/refmatrix_ref.log | head -20 >> $L
ls -la $O >> $L
tail -5 $L
