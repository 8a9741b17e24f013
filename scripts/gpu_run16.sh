#!/bin/bash
# round 2, GPU call 16: one MMA-issuing warp per tile with a token (FA_P4_MMA2) vs the single issuing warp
L=gpurun_out/r02_run16.log
mkdir -p gpurun_out; : > $L
echo "== smoke MMA2 (tiny shapes first: a hang must not eat the box)" >> $L
timeout 120 python scripts/ab_time.py --iters 2 1,512,4,128,0 2,1000,4,128,1 C2 >> $L 2>&1 || { echo "SMOKE FAILED rc=$?" >> $L; tail -5 $L; exit 1; }
echo "== A/B single issuing warp (FA_P4_MMA2=0)" >> $L
FA_B200_LIB=ab/mma1/libfa_b200.so timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a S1k >> $L 2>&1
echo "== A/B MMA2" >> $L
timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a S1k >> $L 2>&1
echo "== parity + fuzz on MMA2" >> $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
echo "== trace MMA2" >> $L
FA_B200_LIB=ab/trace/libfa_b200.so timeout 120 python scripts/trace_fwd.py >> $L 2>&1
tail -5 $L
