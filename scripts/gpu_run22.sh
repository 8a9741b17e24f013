#!/bin/bash
# round 2, GPU call 22: random-shape stress of the shipping library (forward + backward vs the fp32 reference)
L=gpurun_out/r02_run22.log
mkdir -p gpurun_out; : > $L
timeout 500 python scripts/fuzz_shapes.py 300 1 >> $L 2>&1; echo "rc=$?" >> $L
timeout 300 python scripts/fuzz_shapes.py 150 2 >> $L 2>&1; echo "rc=$?" >> $L
grep "FAIL\|done\|rc=" $L | cut -c1-400 | head -40
