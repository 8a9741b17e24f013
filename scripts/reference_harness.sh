#!/bin/bash
# Runs the reference's own benchmark harness command line (/root/reference/benchmark.sh:26-37, staged unmodified in
# baseline/_ref/ by baseline/build_ref.sh) against THIS module for a few points of its grid, then classifies the kernel
# names of the resulting CSVs with the reference's own classifier (utils/plot_kernels.py:95-109 clean_kernel_names).
# Proves that the `--kernel-name ::regex:'(flash_.*|fmha_cutlass.*)'` filter and the substring classifier still pick up
# the four kernel families now that the kernels live in namespace fa100.     (GPU box)
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/gpurun_out/refharness; mkdir -p $OUT
export PYTHONPATH=$ROOT/flash-attention-turing_b200:$ROOT/baseline/_ref
KERNEL_REGEX=$(grep -o "KERNEL_REGEX='[^']*'" $ROOT/baseline/_ref/benchmark.sh | cut -d"'" -f2)
echo "KERNEL_REGEX from the reference's benchmark.sh: $KERNEL_REGEX"
batch_size=4; num_heads=16; num_heads_k=16
for pt in "4096 128 False" "4000 64 True" "1024 128 True"; do
  set -- $pt; seqlen=$1; hdim=$2; is_causal=$3
  timeout 600 ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed \
      --kernel-name-base demangled \
      --kernel-name ::regex:"${KERNEL_REGEX}" \
      --csv python -c "import torch; from test_flash_attn import test_flash_attn_bwd; test_flash_attn_bwd(${batch_size}, ${num_heads}, ${num_heads_k}, ${seqlen}, ${seqlen}, ${hdim}, ${is_causal}, torch.float16)" \
      > "$OUT/${seqlen}_${hdim}_${is_causal}.csv" 2> "$OUT/${seqlen}_${hdim}_${is_causal}.err"
  echo "finished ${seqlen}_${hdim}_${is_causal}.csv rc=$?"
done
python $ROOT/scripts/classify_ncu_csv.py $OUT/*.csv
