#!/bin/bash
# round 2, GPU call 39: fused head_dim-64 backward — reduction split / ring depth variants, then the benchmark.sh grid at head_dim 64
L=gpurun_out/r02_run39.log
mkdir -p gpurun_out; : > $L
for v in flash-attention-turing_b200/flash_attn_turing ab/d64red32 ab/d64red8 ab/d64st3 flash-attention-turing_b200/flash_attn_turing; do
  FA_TAG=$(basename $v) FA_B200_LIB=$v/libfa_b200.so timeout 60 python scripts/ab_time.py --bwd --iters 20 D64a D64c 4,2048,16,64,1 2>&1 | grep "bwd burst" >> $L
done
echo "== sweep bf16 d64" >> $L
PYTHONPATH=flash-attention-turing_b200 timeout 150 python scripts/benchmark_sweep.py --dtype bf16 --hdim 64 > gpurun_out/r02c_sweep_d64_bf16.md 2>> $L
echo "== sweep fp16 d64" >> $L
PYTHONPATH=flash-attention-turing_b200 timeout 150 python scripts/benchmark_sweep.py --dtype fp16 --hdim 64 > gpurun_out/r02c_sweep_d64_fp16.md 2>> $L
cut -c1-200 $L; tail -5 gpurun_out/r02c_sweep_d64_bf16.md
