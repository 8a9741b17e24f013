#!/bin/bash
# round 2, GPU call 13: K/V ring released by the softmax warpgroups
L=gpurun_out/r02_run13.log
mkdir -p gpurun_out; : > $L
echo "== parity subset" >> $L
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3 >> $L
for lib in ab/base/libfa_b200.so flash-attention-turing_b200/flash_attn_turing/libfa_b200.so; do
  echo "== A/B $lib" >> $L
  FA_B200_LIB=$lib timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a S1k >> $L 2>&1
done
echo "== trace" >> $L
FA_B200_LIB=ab/trace/libfa_b200.so FA_TRACE_STEPS=28,34 timeout 120 python scripts/trace_fwd.py 4 4096 >> $L 2>&1
tail -3 $L
