#!/bin/bash
# Build the UNMODIFIED reference kernels (ssiu/flash-attention-turing, /root/reference) for sm_100a
# into baseline/_ref/ (git-ignored, travels to the GPU box with gpurun).  Sources are compiled
# where they lie under /root/reference; nothing is copied into this repo.  The module is named
# flash_attn_turing_ref (via -DTORCH_EXTENSION_NAME) so it can be imported next to ours.
# The reference's legacy mma.sync (HMMA.1688) path is the on-box GPU comparator in bench reports.
set -euo pipefail
REF=${REF:-/root/reference}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
OBJ=${OBJ:-/tmp/fa_ref_obj}
mkdir -p "$OUT" "$OBJ"
PY=${PY:-python}
TORCH_INC=$($PY - <<'PY'
from torch.utils.cpp_extension import include_paths
print(" ".join("-I"+p for p in include_paths()))
PY
)
TORCH_LIB=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'lib'))")
PYINC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
INC="-I$REF/csrc/flash_attn -I$REF/csrc/flash_attn/src -I$REF/csrc/cutlass/include -I$REF/csrc/cutlass/tools/util/include $TORCH_INC -I$PYINC -I/usr/local/cuda/include"
DEFS="-DTORCH_EXTENSION_NAME=flash_attn_turing_ref -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1"
NVFLAGS="-std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -lineinfo -Xcompiler -fPIC"
pids=()
for f in $REF/csrc/flash_attn/src/*.cu; do
  o=$OBJ/$(basename "$f" .cu).o
  [ -f "$o" ] || ( nvcc $NVFLAGS $DEFS $INC -c "$f" -o "$o" ) &
  pids+=($!)
done
[ -f $OBJ/flash_api.o ] || g++ -std=c++17 -O2 -fPIC $DEFS $INC -c $REF/csrc/flash_attn/flash_api.cpp -o $OBJ/flash_api.o &
pids+=($!)
for p in "${pids[@]}"; do wait $p; done
g++ -shared -o "$OUT/flash_attn_turing_ref.so" $OBJ/*.o -L"$TORCH_LIB" -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda \
    -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$TORCH_LIB"
# the reference's own test file is needed on the GPU box to run its acceptance matrix unmodified
cp "$REF/test_flash_attn.py" "$OUT/test_flash_attn.py"
# ... and its benchmark harness (ncu command line + kernel-name classifier) for scripts/reference_harness.sh
cp "$REF/benchmark.sh" "$OUT/benchmark.sh"; cp "$REF/utils/plot_kernels.py" "$OUT/plot_kernels.py"
echo "built $OUT/flash_attn_turing_ref.so"
