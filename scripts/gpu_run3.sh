#!/bin/bash
# round 2, GPU call 3: detailed clock64 trace + ncu source-level capture of the forward
L=gpurun_out/r02_run3.log
mkdir -p gpurun_out; : > $L
echo "== trace d128" >> $L
FA_B200_LIB=ab/trace/libfa_b200.so FA_TRACE_STEPS=24,36 timeout 120 python scripts/trace_fwd.py 4 4096 >> $L 2>&1
echo "== ncu" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_fwd -s 3 -c 1 -o gpurun_out/r02_fwd_p4 python scripts/ab_time.py --iters 2 C2 >> $L 2>&1
ls -la gpurun_out/*.ncu-rep >> $L
tail -3 $L
