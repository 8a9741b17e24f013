"""the ctypes binding of include/fa_b200.h lives in the package (flash_attn_turing/cabi.py); the tests import it from here"""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flash-attention-turing_b200", "flash_attn_turing", "cabi.py")
_spec = importlib.util.spec_from_file_location("fa_b200_cabi", _p)   # by path: works without the compiled _C extension
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
