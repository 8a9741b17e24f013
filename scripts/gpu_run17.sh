#!/bin/bash
# round 2, GPU call 17: MMA2 with batched quarter issue + early token (FA_P4_TOK_LEFT 0 / 3 / 5)
L=gpurun_out/r02_run17.log
mkdir -p gpurun_out; : > $L
echo "== smoke" >> $L
timeout 120 python scripts/ab_time.py --iters 2 1,512,4,128,0 2,1000,4,128,1 C2 >> $L 2>&1 || { echo "SMOKE FAILED rc=$?" >> $L; tail -5 $L; exit 1; }
for v in ab/mma1 ab/tok0 ab/tok5 flash-attention-turing_b200/flash_attn_turing; do
  echo "== A/B $v" >> $L
  FA_B200_LIB=$v/libfa_b200.so timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a >> $L 2>&1
done
echo "== parity + fuzz (default = TOK_LEFT 3)" >> $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
echo "== trace" >> $L
FA_B200_LIB=ab/trace/libfa_b200.so timeout 120 python scripts/trace_fwd.py >> $L 2>&1
grep "^AB\|passed\|failed" $L | cut -c1-200
