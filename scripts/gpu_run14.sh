#!/bin/bash
# round 2, GPU call 14: paired causal schedule
L=gpurun_out/r02_run14.log
mkdir -p gpurun_out; : > $L
echo "== pytest gpu (all)" >> $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
for lib in ab/base/libfa_b200.so flash-attention-turing_b200/flash_attn_turing/libfa_b200.so; do
  echo "== A/B $lib" >> $L
  FA_B200_LIB=$lib timeout 300 python scripts/ab_time.py --sustain 1 C3 C2c 4,16384,16,128,1 4,8192,16,128,1 4,2048,16,128,1 D64c C2 >> $L 2>&1
done
tail -3 $L
