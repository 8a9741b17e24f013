#!/bin/bash
# round 2, GPU call 4: straight-line mask copies; timing experiment without the per-step max pass
L=gpurun_out/r02_run4.log
mkdir -p gpurun_out; : > $L
for lib in flash-attention-turing_b200/flash_attn_turing/libfa_b200.so ab/nomax/libfa_b200.so; do
  echo "== A/B $lib" >> $L
  FA_B200_LIB=$lib timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a S1k >> $L 2>&1
  FA_B200_LIB=$lib FA_B200_EMU=3 FA_TAG="$lib EMU=3" timeout 300 python scripts/ab_time.py C2 C3 D64a >> $L 2>&1
  FA_B200_LIB=$lib FA_B200_EMU=0 FA_TAG="$lib EMU=0" timeout 300 python scripts/ab_time.py C2 C3 D64a >> $L 2>&1
done
echo "== trace nomax" >> $L
FA_B200_LIB=ab/nomaxtrace/libfa_b200.so FA_TRACE_STEPS=24,34 timeout 120 python scripts/trace_fwd.py 4 4096 >> $L 2>&1
tail -3 $L
