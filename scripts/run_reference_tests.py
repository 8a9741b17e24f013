"""Run the reference's OWN test file (baseline/_ref/test_flash_attn.py, staged by baseline/build_ref.sh) against this
module or against the reference's own kernels rebuilt for sm_100a, all parametrized cases (2 528 dense + 2 560 varlen at
stride 1), and summarise.  Inputs are seeded per case so that both implementations see identical tensors.
usage: run_reference_tests.py [stride [ours|ref [default|math]]]   (GPU box)"""
import importlib.util, itertools, os, sys, io, contextlib, collections, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200"))
import torch
spec = importlib.util.spec_from_file_location("ref_test_flash_attn", os.path.join(ROOT, "baseline", "_ref", "test_flash_attn.py"))
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 1
impl = sys.argv[2] if len(sys.argv) > 2 else "ours"          # ours | ref  (ref = the reference's own kernels rebuilt for sm_100a)
backend = sys.argv[3] if len(sys.argv) > 3 else "default"     # default | math  (torch SDPA backend used by the reference's oracle)
if impl == "ref":
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import flash_attn_turing_ref as _ref
    for name in ("fwd", "bwd", "varlen_fwd", "varlen_bwd"):
        setattr(mod, name, getattr(_ref, name))
from torch.nn.attention import SDPBackend, sdpa_kernel
ctx = (lambda: sdpa_kernel(SDPBackend.MATH)) if backend == "math" else contextlib.nullcontext

def params_of(fn):
    marks = [m for m in fn.pytestmark if m.name == "parametrize"]
    names, values = [], []
    for m in reversed(marks):
        n = [x.strip() for x in m.args[0].split(",")]
        names.append(n); values.append(m.args[1])
    for combo in itertools.product(*values):
        kw = {}
        for n, v in zip(names, combo):
            if len(n) == 1: kw[n[0]] = v
            else: kw.update(dict(zip(n, v)))
        yield kw

for fn in (mod.test_flash_attn_bwd, mod.test_flash_attn_bwd_varlen):
    cases = list(params_of(fn))[::stride]
    res = collections.Counter(); fails = []
    t0 = time.time()
    for ci, kw in enumerate(cases):
        torch.manual_seed(1234 + ci)      # the reference draws unseeded from the global RNG: same inputs for both implementations
        try:
            with contextlib.redirect_stdout(io.StringIO()), ctx():
                fn(**kw)
            res["pass"] += 1
        except AssertionError as e:
            res["fail"] += 1; fails.append((kw, str(e)[:120]))
        except Exception as e:  # noqa
            res["error"] += 1; fails.append((kw, "EXC " + repr(e)[:200]))
    print(f"REFTEST impl={impl} sdpa={backend} {fn.__name__}: {len(cases)} cases {dict(res)} in {time.time()-t0:.0f}s", flush=True)
    kinds = collections.Counter(f[1].split("=")[0] for f in fails)
    print("REFTEST failure kinds:", dict(kinds))
    for kw, msg in fails[:6]:
        print("REFTEST FAIL", {k: (str(v) if not isinstance(v, (int, bool)) else v) for k, v in kw.items()}, msg)
