#!/bin/bash
# round 2, GPU call 42: more seeds of the random-shape stress on the shipping build (fused head_dim-64 backward by default)
L=gpurun_out/r02_run42.log
mkdir -p gpurun_out; : > $L
for seed in 7 1 2 3 4; do
  timeout 100 python scripts/fuzz_shapes.py 200 $seed 2>&1 | tail -2 >> $L
done
cat $L | cut -c1-200
