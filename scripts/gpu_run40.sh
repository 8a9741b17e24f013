#!/bin/bash
# round 2, GPU call 40: fused head_dim-64 backward, Q/dO ring of 8 stages against 6
L=gpurun_out/r02_run40.log
mkdir -p gpurun_out; : > $L
for v in flash-attention-turing_b200/flash_attn_turing ab/d64st8 flash-attention-turing_b200/flash_attn_turing ab/d64st8; do
  FA_TAG=$(basename $v) FA_B200_LIB=$v/libfa_b200.so timeout 60 python scripts/ab_time.py --bwd --iters 20 D64a D64c 4,2048,16,64,1 4,16384,16,64,0 2>&1 | grep "bwd burst" >> $L
done
cut -c1-200 $L
