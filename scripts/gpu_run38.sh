#!/bin/bash
# round 2, GPU call 38: ab_time's one-head sanity check reads 0.3 for B4h16d64 after D64a, D64c with --sustain: ours or torch's autograd?
L=gpurun_out/r02_run38.log
mkdir -p gpurun_out; : > $L
timeout 150 python scripts/ab_time.py --bwd --sustain 0.5 D64a D64c B4h16d64 >> $L 2>&1
grep "bwd burst\|rror" $L | cut -c1-330
