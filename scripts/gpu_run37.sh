#!/bin/bash
# round 2, GPU call 37: is the forward bit-reproducible under a sustained loop?
L=gpurun_out/r02_run37.log
mkdir -p gpurun_out; : > $L
timeout 300 python scripts/diag_repeat.py --seconds 1.5 4,4096,32,64,0 4,8192,32,64,1 4,4096,16,64,0 4,4096,32,128,0 4,8192,32,128,1 >> $L 2>&1
grep "REPEAT\|rror" $L | cut -c1-250
