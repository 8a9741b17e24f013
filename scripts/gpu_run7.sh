#!/bin/bash
L=gpurun_out/r02_run7.log
mkdir -p gpurun_out; : > $L
echo "== trace (epilogue rows 6/7)" >> $L
FA_B200_LIB=ab/trace/libfa_b200.so FA_TRACE_STEPS=30,34 timeout 120 python scripts/trace_fwd.py 4 4096 >> $L 2>&1
FA_B200_LIB=ab/trace/libfa_b200.so FA_TRACE_STEPS=6,10 timeout 120 python scripts/trace_fwd.py 16 1024 >> $L 2>&1
tail -3 $L
