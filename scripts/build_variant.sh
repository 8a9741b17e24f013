#!/bin/bash
# build_variant.sh NAME "EXTRA NVCC FLAGS": builds libfa_b200.so with extra -D flags into ab/NAME/ (A/B runs: FA_B200_LIB=ab/NAME/libfa_b200.so)
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/flash-attention-turing_b200/csrc/flash_attn/src
OUT=$ROOT/ab/$NAME; mkdir -p $OUT/obj
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --use_fast_math -Xptxas -v -Xcompiler -fPIC $*"
for f in fa_api flash_fwd_sm100 flash_fwd_p4_sm100 flash_bwd_sm100 flash_bwd_tc_sm100; do
  nvcc $FLAGS -c $SRC/$f.cu -o $OUT/obj/$f.o 2> $OUT/obj/$f.log &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libfa_b200.so $OUT/obj/*.o -cudart shared
grep -h "spill" $OUT/obj/flash_fwd_p4_sm100.log | sort | uniq -c | head -4
echo "built $OUT/libfa_b200.so"
