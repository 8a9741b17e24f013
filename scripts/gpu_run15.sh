#!/bin/bash
# round 2, GPU call 15: finer MMA-warp trace; A/B of the issue-by-readiness MMA loop (FA_P4_MMA_POLL)
L=gpurun_out/r02_run15.log
mkdir -p gpurun_out; : > $L
echo "== trace (default build, finer MMA-warp stamps)" >> $L
FA_B200_LIB=ab/trace/libfa_b200.so timeout 120 python scripts/trace_fwd.py >> $L 2>&1
echo "== A/B default" >> $L
FA_B200_LIB=flash-attention-turing_b200/flash_attn_turing/libfa_b200.so timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a >> $L 2>&1
echo "== A/B poll" >> $L
FA_B200_LIB=ab/poll/libfa_b200.so timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a >> $L 2>&1
echo "== parity on poll" >> $L
FA_B200_LIB=ab/poll/libfa_b200.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3 >> $L
echo "== trace poll" >> $L
FA_B200_LIB=ab/polltrace/libfa_b200.so timeout 120 python scripts/trace_fwd.py >> $L 2>&1
tail -5 $L
