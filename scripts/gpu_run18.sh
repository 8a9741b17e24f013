#!/bin/bash
# round 2, GPU call 18: MMA2, token after the last MMA; batched quarter issue with / without the pre-waited first quarter
L=gpurun_out/r02_run18.log
mkdir -p gpurun_out; : > $L
echo "== smoke" >> $L
timeout 120 python scripts/ab_time.py --iters 2 1,512,4,128,0 2,1000,4,128,1 C2 >> $L 2>&1 || { echo "SMOKE FAILED rc=$?" >> $L; tail -5 $L; exit 1; }
for v in ab/mma1 ab/nopre flash-attention-turing_b200/flash_attn_turing ab/mma1 ab/nopre flash-attention-turing_b200/flash_attn_turing; do
  echo "== A/B $v" >> $L
  FA_B200_LIB=$v/libfa_b200.so timeout 120 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a >> $L 2>&1
done
echo "== parity + fuzz (default)" >> $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
grep "^AB\|passed\|failed" $L | cut -c1-60,140-240
