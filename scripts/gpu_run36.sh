#!/bin/bash
# round 2, GPU call 36: does the order of shapes inside one process matter? (ab_time's sanity check read 0.3 for B4h16d64 after D64a, D64c)
L=gpurun_out/r02_run36.log
mkdir -p gpurun_out; : > $L
timeout 200 python scripts/diag_shape.py 4,4096,32,64,0 4,8192,32,64,1 4,4096,16,64,0 >> $L 2>&1
timeout 150 python scripts/ab_time.py --bwd --sustain 0.5 D64a D64c B4h16d64 >> $L 2>&1
timeout 150 python scripts/ab_time.py --bwd D64c B4h16d64 >> $L 2>&1
grep "SHAPE\|bwd burst\|rror" $L | cut -c1-250
