import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(ROOT, "flash-attention-turing_b200")
for p in (ROOT, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _bf16_bits_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def load_golden(path):
    """-> dict with q,k,v,dout as float32 (exact 16-bit values) + reference outputs + meta"""
    z = np.load(path)
    dtype = str(z["dtype"])
    out = {"name": os.path.basename(path)[:-4], "dtype": dtype, "causal": bool(z["causal"])}
    for key in ("q", "k", "v", "dout"):
        a = z[key]
        out[key] = _bf16_bits_to_f32(a) if a.dtype == np.uint16 else a.astype(np.float32)
    for key in ("out", "dq", "dk", "dv", "cu_q", "cu_k"):
        if key in z.files:
            out[key] = z[key]
    if "max_sq" in z.files:
        out["max_sq"], out["max_sk"] = int(z["max_sq"]), int(z["max_sk"])
    return out


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "_*.npz")))


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle
