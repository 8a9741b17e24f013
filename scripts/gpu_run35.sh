#!/bin/bash
# round 2, GPU call 35: per-head errors at the benchmark.sh grid point b4 s4096 h16 d64 (ab_time's one-head sanity check read 0.3 there)
L=gpurun_out/r02_run35.log
mkdir -p gpurun_out; : > $L
timeout 200 python scripts/diag_shape.py 4,4096,16,64,0 4,4096,32,64,0 4,4096,16,128,0 2,2048,16,64,0 4,4096,16,64,1 >> $L 2>&1
FA_B200_BWD_D64=det timeout 100 python scripts/diag_shape.py 4,4096,16,64,0 >> $L 2>&1
timeout 100 python scripts/ab_time.py --bwd --iters 3 B4h16d64 B4h16 >> $L 2>&1
grep "SHAPE\|bwd\|rror" $L | cut -c1-250
