"""CPU tests of the drop-in boundary: libfa_b200.so loads, exports every symbol include/fa_b200.h declares, and its
argument validation (which runs before any device work) reports errors the way the header documents."""
import ctypes

import pytest

import cabi


def test_library_loads_and_exports_every_declared_symbol():
    lib = cabi.load()
    names = cabi.declared_symbols()
    assert {"fa_b200_fwd", "fa_b200_bwd", "fa_b200_bwd_workspace_bytes", "fa_b200_last_error",
            "fa_b200_last_launch_count", "fa_b200_abi_version"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fa_b200.h but not exported"
    assert lib.fa_b200_abi_version() == 1


def _params(**kw):
    p = cabi.FwdParams()
    # non-null fake pointers: validation must reject before touching them
    p.q = p.k = p.v = p.o = p.lse = 0x1000
    p.b, p.seqlen_q, p.seqlen_k, p.h, p.h_k, p.d = 1, 128, 128, 4, 2, 128
    p.dtype, p.is_causal = cabi.FA_DTYPE_BF16, 0
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.mark.parametrize("kw,msg", [
    (dict(d=96), "head_dim"),
    (dict(h=6, h_k=4), "divisible"),
    (dict(dtype=7), "dtype"),
    (dict(q=None), "null"),
    (dict(cu_seqlens_q=0x2000), "cu_seqlens"),
    (dict(q=0x1008), "aligned"),
    (dict(seqlen_q=-1), "sizes"),
])
def test_invalid_arguments_are_rejected_without_a_device(kw, msg):
    lib = cabi.load()
    p = _params(**kw)
    rc = lib.fa_b200_fwd(ctypes.byref(p), None)
    assert rc == cabi.FA_ERR_INVALID_ARG
    assert msg in lib.fa_b200_last_error().decode()


def test_null_params():
    lib = cabi.load()
    assert lib.fa_b200_fwd(None, None) == cabi.FA_ERR_INVALID_ARG
    assert lib.fa_b200_bwd(None, None) == cabi.FA_ERR_INVALID_ARG


def test_bwd_validates_gradient_pointers():
    lib = cabi.load()
    p = cabi.BwdParams()
    p.fwd = _params()
    rc = lib.fa_b200_bwd(ctypes.byref(p), None)
    assert rc == cabi.FA_ERR_INVALID_ARG and "dout" in lib.fa_b200_last_error().decode()


def test_bwd_workspace_bytes_is_the_fused_backward_accumulator():
    """host-only: the fp32 dQ accumulator [b][h][ceil64(seqlen_q)][d] of the fused backward, head_dim 128 and 64"""
    import os
    lib = cabi.load()
    if os.environ.get("FA_B200_BWD", "fused")[0] != "f" or os.environ.get("FA_B200_BWD_D64", "fused")[0] != "f":
        pytest.skip("FA_B200_BWD / FA_B200_BWD_D64 select a non-fused backward")
    p = _params(b=3, seqlen_q=130, h=6, h_k=2, d=128)
    assert lib.fa_b200_bwd_workspace_bytes(ctypes.byref(p)) == 3 * 6 * 192 * 128 * 4
    p = _params(b=3, seqlen_q=130, h=6, h_k=2, d=64)
    assert lib.fa_b200_bwd_workspace_bytes(ctypes.byref(p)) == 3 * 6 * 192 * 64 * 4
    assert lib.fa_b200_bwd_workspace_bytes(None) == 0


def test_struct_layout_matches_header():
    # 7 pointers + 8 int64 + 2 int32 = 128 bytes; bwd adds 6 pointers
    assert ctypes.sizeof(cabi.FwdParams) == 7 * 8 + 8 * 8 + 2 * 4
    assert ctypes.sizeof(cabi.BwdParams) == ctypes.sizeof(cabi.FwdParams) + 6 * 8


def test_python_package_fails_loudly_without_extension(monkeypatch, tmp_path):
    """the product path has no fallback: a missing compiled extension is an ImportError, not a silent CPU path"""
    import importlib.util
    import os
    import shutil
    import sys
    src = os.path.join(cabi.ROOT, "flash-attention-turing_b200", "flash_attn_turing", "__init__.py")
    pkg = tmp_path / "flash_attn_turing"
    pkg.mkdir()
    shutil.copy(src, pkg / "__init__.py")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "flash_attn_turing" or k.startswith("flash_attn_turing.")}
    try:
        spec = importlib.util.spec_from_file_location("flash_attn_turing", str(pkg / "__init__.py"),
                                                      submodule_search_locations=[str(pkg)])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["flash_attn_turing"] = mod
        with pytest.raises(ImportError, match="no CPU / PyTorch fallback"):
            spec.loader.exec_module(mod)
    finally:
        for k in list(sys.modules):
            if k == "flash_attn_turing" or k.startswith("flash_attn_turing."):
                del sys.modules[k]
        sys.modules.update(saved)


def test_library_sass_is_tcgen05_tma_code_without_legacy_mma():
    """host-only (cuobjdump): the shipped kernels are Blackwell-native — tcgen05.mma (UTCHMMA), TMEM loads / stores (LDTM / STTM),
    TMA tensor loads / stores (UTMALDG / UTMASTG), bulk reductions (UBLKRED) — and contain no legacy mma.sync (HMMA); the forward
    and the fused backward exist for head_dim 64 and 128 in both 16-bit types (profiles/r02c_sass_counts.txt is this, per kernel)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", cabi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    count = {m: sass.count(m) for m in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKRED")}
    assert all(v > 0 for v in count.values()), count
    assert " HMMA." not in sass and "HMMA.16" not in sass, "legacy mma.sync code in the library"
    funcs = [ln.split("Function :")[1].strip() for ln in sass.splitlines() if "Function :" in ln]
    for d in (64, 128):
        for b in (0, 1):
            assert any(f"flash_fwd_kernel_sm100_p4ILi{d}ELb{b}E" in f for f in funcs), (d, b)
            assert any(f"flash_bwd_dk_dv_kernel_sm100_fusedILi{d}ELb{b}E" in f for f in funcs), (d, b)
