#!/bin/bash
# round 2, GPU call 21: token through a hardware named barrier / first P quarter waited for before the token
L=gpurun_out/r02_run21.log
mkdir -p gpurun_out; : > $L
for v in ab/named ab/prew ab/namedprew; do
  FA_B200_LIB=$v/libfa_b200.so timeout 100 python scripts/ab_time.py --iters 2 1,512,4,128,0 2,1000,4,128,1 3,700,6,64,1 >> $L 2>&1 || echo "SMOKE $v FAILED rc=$?" >> $L
done
for v in flash-attention-turing_b200/flash_attn_turing ab/named ab/prew ab/namedprew flash-attention-turing_b200/flash_attn_turing ab/named ab/prew ab/namedprew; do
  echo "== A/B $v" >> $L
  FA_B200_LIB=$v/libfa_b200.so timeout 120 python scripts/ab_time.py --sustain 1 C2 C3 D64a >> $L 2>&1
done
grep "^AB\|passed\|failed\|FAILED" $L | grep -v "n=2:" | sed 's/err vs SDPA [^|]*|//' | cut -c1-200
