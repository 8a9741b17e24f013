"""Random-shape stress through the C ABI (scripts/fuzz_shapes.py): tiny items, one-tile items, query tiles with no visible key,
unequal block counts of the two query tiles, GQA, head_dim 64 / 128, fp16 / bf16, score scales that trigger the retry pass;
forward always, backward on 40 % of the shapes; against the fp32 reference with the stated gates (tests/gpu_ref.py).  Runs in
a subprocess under a timeout: a hang would otherwise take the whole suite with it.  (All three seeds hung the first retry scheme
of round 2 — a retry list whose length changed while pass 1 was running; FUZZ_VERBOSE=1 prints every
shape before it runs.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12, 7])
def test_random_shapes(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_shapes.py"), "200", str(seed)],
                       capture_output=True, text=True, timeout=300)
    tail = "\n".join(r.stdout.strip().splitlines()[-15:])
    assert r.returncode == 0, f"fuzz_shapes failed:\n{tail}\n{r.stderr[-2000:]}"
    assert "200 shapes, 0 failures" in r.stdout, tail
