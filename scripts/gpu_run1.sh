#!/bin/bash
# round 2, GPU call 1: new forward loop (max-first, quartered P release) — correctness, A/B against the round-1 library, trace
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r02_run1.log 2>&1
echo "== quick sanity" >> gpurun_out/r02_run1.log
timeout 300 python scripts/ab_time.py S1k 2,300,4,128,1 2,333,4,64,1 >> gpurun_out/r02_run1.log 2>&1 || echo "SANITY FAILED rc=$?" >> gpurun_out/r02_run1.log
echo "== pytest gpu (parity + api)" >> gpurun_out/r02_run1.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> gpurun_out/r02_run1.log
for lib in ab/r01/libfa_b200.so flash-attention-turing_b200/flash_attn_turing/libfa_b200.so; do
  echo "== A/B $lib" >> gpurun_out/r02_run1.log
  FA_B200_LIB=$lib timeout 300 python scripts/ab_time.py --sustain 2 C2 C3 C4 D64a D64c S1k >> gpurun_out/r02_run1.log 2>&1
done
for emu in 0 4 3 2; do
  echo "== new lib FA_B200_EMU=$emu" >> gpurun_out/r02_run1.log
  FA_B200_EMU=$emu timeout 200 python scripts/ab_time.py C2 C3 D64a >> gpurun_out/r02_run1.log 2>&1
done
echo "== torch SDPA" >> gpurun_out/r02_run1.log
timeout 200 python scripts/ab_time.py --sdpa C2 C3 C4 D64a 2>&1 | grep torch-SDPA >> gpurun_out/r02_run1.log
echo "== trace" >> gpurun_out/r02_run1.log
FA_B200_LIB=ab/trace/libfa_b200.so timeout 120 python scripts/trace_fwd.py 4 4096 >> gpurun_out/r02_run1.log 2>&1
tail -5 gpurun_out/r02_run1.log
