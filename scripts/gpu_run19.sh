#!/bin/bash
# round 2, GPU call 19: MMA2 (late token, in-order quarters) vs single hand-over of P vs the single issuing warp; EMU variants on MMA2
L=gpurun_out/r02_run19.log
mkdir -p gpurun_out; : > $L
echo "== smoke" >> $L
timeout 120 python scripts/ab_time.py --iters 2 1,512,4,128,0 2,1000,4,128,1 C2 >> $L 2>&1 || { echo "SMOKE FAILED rc=$?" >> $L; tail -5 $L; exit 1; }
FA_B200_LIB=ab/onehand/libfa_b200.so timeout 120 python scripts/ab_time.py --iters 2 1,512,4,128,0 2,1000,4,128,1 >> $L 2>&1 || echo "SMOKE onehand FAILED rc=$?" >> $L
for v in ab/mma1 flash-attention-turing_b200/flash_attn_turing ab/onehand ab/mma1 flash-attention-turing_b200/flash_attn_turing ab/onehand; do
  echo "== A/B $v" >> $L
  FA_B200_LIB=$v/libfa_b200.so timeout 120 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a >> $L 2>&1
done
for e in 1 3 4; do
  echo "== EMU=$e (default lib)" >> $L
  FA_B200_EMU=$e FA_TAG=EMU$e timeout 120 python scripts/ab_time.py --sustain 1 C2 C3 D64a >> $L 2>&1
done
grep "^AB\|passed\|failed\|FAILED" $L | cut -c1-60,140-240
