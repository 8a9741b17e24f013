"""CPU model check of the forward kernel's online softmax (scripts/model_p4_softmax.py): float32 emulation of the exact
first step, the speculative steps with the late-agreed reference (16-bit sum exchange, power-of-two shifts), the overflow flag
that sends an item to the exact redo, and the polynomial exp2 with its raw-bits exponent add.  The B200 found a range hole
in this kind of logic in round 1 (profiles/r01s2_rescale_bug.log: without the upper clamp the polynomial wraps around and
the overflow goes unnoticed); the model reproduces it with the clamp switched off and must be clean with it on, for the
bf16 and the fp16 overflow limits."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "model_p4_softmax.py")


def _run(*args):
    r = subprocess.run([sys.executable, SCRIPT, *args], capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout.strip().splitlines()[-1]


@pytest.mark.parametrize("args", [(), ("--fp16",)])
def test_model_is_clean_with_the_clamped_polynomial(args):
    rc, last = _run(*args)
    assert rc == 0 and " 0 bad" in last and "0 of 500 benign rows flagged" in last, last


def test_model_reproduces_the_wrap_around_without_the_clamp():
    rc, last = _run("--no-clamp")
    assert rc != 0 and " 0 bad" not in last, last
