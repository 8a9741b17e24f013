"""Build surface of the B200-native flash_attn_turing.

Mirrors the reference's setup.py (/root/reference/setup.py:20-49: one torch extension named flash_attn_turing built
from csrc/flash_attn/) with two changes: the CUDA kernels live in a torch-free C-ABI library (libfa_b200.so,
include/fa_b200.h) compiled by plain nvcc for sm_100a, and the pybind layer (csrc/flash_attn/flash_api.cpp) is a
host-only C++ extension linked against it.  No CUTLASS include dirs, no `git submodule update`.

    python setup.py build_ext --inplace      # in-tree build (what __graft_entry__.build() runs)
    pip install --no-build-isolation .       # the reference's install.sh / benchmark.sh path
"""
import os
import subprocess
import sys

from setuptools import setup
from torch.utils.cpp_extension import BuildExtension, CppExtension

this_dir = os.path.dirname(os.path.abspath(__file__))
pkg_root = os.path.join(this_dir, "flash-attention-turing_b200")
pkg_dir = os.path.join(pkg_root, "flash_attn_turing")

# 1) the kernel library: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (see the Makefile)
subprocess.run(["make", "-C", pkg_root, "-j", str(os.cpu_count() or 4)], check=True)

cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")

setup(
    name="flash_attn_turing",
    version="0.1.0+b200",
    package_dir={"": "flash-attention-turing_b200"},
    packages=["flash_attn_turing"],
    package_data={"flash_attn_turing": ["libfa_b200.so"]},
    ext_modules=[
        CppExtension(
            name="flash_attn_turing._C",
            sources=["flash-attention-turing_b200/csrc/flash_attn/flash_api.cpp"],
            include_dirs=[os.path.join(this_dir, "include"), os.path.join(cuda_home, "include")],
            library_dirs=[pkg_dir],
            libraries=["fa_b200", "c10_cuda", "torch_cuda"],
            runtime_library_dirs=["$ORIGIN"],
            extra_compile_args=["-O2", "-g0", "-std=c++17"],
        )
    ],
    install_requires=["torch"],
    cmdclass={"build_ext": BuildExtension},
)
