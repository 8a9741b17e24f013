"""CPU model check of the p4 forward's online softmax (scripts/model_p4_softmax.py): float32 emulation of the lazily
updated reference max, the speculative exponentials, the tile-sum vote and the polynomial exp2 with its raw-bits exponent
add.  The B200 run found a range hole in exactly this logic (profiles/r01s2_rescale_bug.log); the model reproduces it with
the clamp switched off and must be clean with it on."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "model_p4_softmax.py")


def _run(*args):
    r = subprocess.run([sys.executable, SCRIPT, *args], capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout.strip().splitlines()[-1]


def test_model_is_clean_with_the_clamped_polynomial():
    rc, last = _run()
    assert rc == 0 and " 0 bad" in last, last


def test_model_reproduces_the_wrap_around_without_the_clamp():
    rc, last = _run("--no-clamp")
    assert rc != 0 and " 0 bad" not in last, last
