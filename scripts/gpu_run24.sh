#!/bin/bash
# round 2, GPU call 24: epilogue that stores O from registers (32-byte global stores) vs staging tile + TMA store
L=gpurun_out/r02_run24.log
mkdir -p gpurun_out; : > $L
timeout 100 python scripts/ab_time.py --iters 2 1,512,4,128,0 2,1000,4,128,1 3,700,6,64,1 >> $L 2>&1 || { echo "SMOKE FAILED rc=$?" >> $L; tail -5 $L; exit 1; }
for v in ab/tmastore flash-attention-turing_b200/flash_attn_turing ab/tmastore flash-attention-turing_b200/flash_attn_turing; do
  echo "== A/B $v" >> $L
  FA_B200_LIB=$v/libfa_b200.so timeout 200 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a S1k S512 4,2048,16,128,1 4,1024,16,128,1 >> $L 2>&1
done
echo "== tests" >> $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
echo "== trace" >> $L
FA_B200_LIB=ab/trace/libfa_b200.so timeout 120 python scripts/trace_fwd.py >> $L 2>&1
grep "^AB\|passed\|failed\|FAILED" $L | grep -v "n=2:" | sed 's/err vs SDPA [^|]*|//; s/flash-attention-turing_b200\/flash_attn_turing/NEW/' | cut -c1-200
