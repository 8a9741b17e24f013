#!/bin/bash
# round 2, GPU call 26: fused backward at head_dim 64 (M = 64 dQ^T) — correctness, A/B against the two deterministic
# kernels, variants (reduction split, polynomial exponentials), clock64 trace
L=gpurun_out/r02_run26.log
mkdir -p gpurun_out; : > $L
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm')" >> $L 2>&1
export FA_B200_BWD_D64=fused
echo "== smoke (fused d64)" >> $L
timeout 150 python scripts/ab_time.py --bwd --iters 2 1,512,4,64,0 2,1000,4,64,1 3,700,6,64,1 1,256,2,64,0 2,192,4,128,1 >> $L 2>&1 || { echo "SMOKE FAILED rc=$?" >> $L; tail -8 $L; exit 1; }
echo "== tests with FA_B200_BWD_D64=fused" >> $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $L
echo "== random shapes (fused d64)" >> $L
timeout 300 python scripts/fuzz_shapes.py 150 7 2>&1 | tail -4 >> $L
SH="D64a D64c B4h16d64 4,1024,16,64,0 4,2048,16,64,1 4,16384,16,64,1"
echo "== A/B det" >> $L
FA_B200_BWD_D64=det FA_TAG=det timeout 300 python scripts/ab_time.py --bwd --sustain 0.5 $SH >> $L 2>&1
echo "== A/B fused" >> $L
FA_TAG=fused timeout 300 python scripts/ab_time.py --bwd --sustain 0.5 $SH >> $L 2>&1
for v in d64red0 d64red32 d64emu2 d64emu4; do
  echo "== A/B $v" >> $L
  FA_TAG=$v FA_B200_LIB=ab/$v/libfa_b200.so timeout 300 python scripts/ab_time.py --bwd --sustain 0.5 D64a D64c 4,2048,16,64,1 >> $L 2>&1
done
echo "== trace d64 fused" >> $L
FA_B200_LIB=ab/trace/libfa_b200.so timeout 120 python scripts/trace_bwd.py 4 4096 64 >> $L 2>&1
grep "bwd\|passed\|failed\|FAILED\|clean\|mismatch" $L | grep -v "n=2:" | cut -c1-220
