#!/bin/bash
# round 2, GPU call 10: late-agreed reference with the four-slot K/V ring
L=gpurun_out/r02_run10.log
mkdir -p gpurun_out; : > $L
echo "== pytest gpu (parity+fuzz)" >> $L
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_api_gpu.py -m gpu -x -q 2>&1 | tail -6 >> $L
echo "== A/B default (speculative)" >> $L
timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a S1k S512 >> $L 2>&1
FA_B200_EMU=3 FA_TAG="EMU=3" timeout 300 python scripts/ab_time.py C2 C3 D64a >> $L 2>&1
FA_B200_EMU=0 FA_TAG="EMU=0" timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 D64a S1k >> $L 2>&1
echo "== A/B exact (FA_B200_FWD_EXACT=1)" >> $L
FA_B200_FWD_EXACT=1 FA_TAG="EXACT" timeout 300 python scripts/ab_time.py --sustain 1 C2 C3 C4 D64a S1k >> $L 2>&1
echo "== fp16 / sdpa" >> $L
timeout 300 python scripts/ab_time.py --dtype fp16 --sdpa C2 C3 >> $L 2>&1
echo "== trace" >> $L
FA_B200_LIB=ab/trace/libfa_b200.so FA_TRACE_STEPS=28,36 timeout 120 python scripts/trace_fwd.py 4 4096 >> $L 2>&1
tail -3 $L
