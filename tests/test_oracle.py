"""CPU tests: pin the oracle (oracle/attn_oracle.c) against golden vectors produced by the REFERENCE's own Python
reference functions (tests/golden/make_golden.py imports /root/reference/test_flash_attn.py), and against torch."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden

# the golden outputs are float32 torch results; the oracle accumulates in double -> agreement to fp32 round-off
ATOL = 2e-5


@pytest.mark.parametrize("path", golden_files("dense"), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_python_dense(oracle_mod, path):
    g = load_golden(path)
    o, lse = oracle_mod.attention_fwd(g["q"], g["k"], g["v"], g["causal"])
    np.testing.assert_allclose(o, g["out"], atol=ATOL, rtol=1e-4)
    if "dq" in g:
        dq, dk, dv = oracle_mod.attention_bwd(g["q"], g["k"], g["v"], o, lse, g["dout"], g["causal"])
        np.testing.assert_allclose(dq, g["dq"], atol=5e-5, rtol=1e-4)
        np.testing.assert_allclose(dk, g["dk"], atol=5e-5, rtol=1e-4)
        np.testing.assert_allclose(dv, g["dv"], atol=5e-5, rtol=1e-4)


@pytest.mark.parametrize("path", golden_files("varlen"), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_python_varlen(oracle_mod, path):
    g = load_golden(path)
    kw = dict(cu_q=g["cu_q"], cu_k=g["cu_k"], max_sq=g["max_sq"], max_sk=g["max_sk"])
    o, lse = oracle_mod.attention_fwd(g["q"], g["k"], g["v"], g["causal"], **kw)
    np.testing.assert_allclose(o, g["out"], atol=ATOL, rtol=1e-4)
    dq, dk, dv = oracle_mod.attention_bwd(g["q"], g["k"], g["v"], o, lse, g["dout"], g["causal"], **kw)
    np.testing.assert_allclose(dq, g["dq"], atol=5e-5, rtol=1e-4)
    np.testing.assert_allclose(dk, g["dk"], atol=5e-5, rtol=1e-4)
    np.testing.assert_allclose(dv, g["dv"], atol=5e-5, rtol=1e-4)
    # LSE padding beyond each sequence stays zero (flash_api.cpp:352 allocates l with torch::zeros)
    lens = np.diff(g["cu_q"])
    for b, n in enumerate(lens):
        assert np.all(lse[b, :, n:] == 0)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("sq,sk", [(1, 1), (5, 9), (9, 5), (64, 64), (33, 130)])
def test_oracle_lse_and_empty_rows(oracle_mod, sq, sk, causal):
    """LSE is the natural-log LSE of the scaled scores and 0 for rows with no visible key
    (flash_fwd_kernel.h:766-785); such rows give O = 0 (:717-730)."""
    torch.manual_seed(sq * 1000 + sk)
    b, h, hk, d = 2, 4, 2, 64
    q, k, v = torch.randn(b, sq, h, d), torch.randn(b, sk, hk, d), torch.randn(b, sk, hk, d)
    o, lse = oracle_mod.attention_fwd(q.numpy(), k.numpy(), v.numpy(), causal)
    kk = k.repeat_interleave(h // hk, dim=2)
    s = torch.einsum("bqhd,bkhd->bhqk", q.double(), kk.double()) / math.sqrt(d)
    if causal:
        mask = torch.tril(torch.ones(sq, sk, dtype=torch.bool), diagonal=sk - sq)
        s = s.masked_fill(~mask, float("-inf"))
    ref = torch.logsumexp(s, dim=-1)
    empty = torch.isinf(ref)
    ref = torch.where(empty, torch.zeros_like(ref), ref)
    np.testing.assert_allclose(lse, ref.numpy(), atol=1e-5)
    if empty.any():
        o_t = torch.from_numpy(o).permute(0, 2, 1, 3)  # b h q d
        assert torch.all(o_t[empty] == 0)


def test_oracle_fast_variant_agrees(oracle_mod):
    """the float32 variant timed by bench.py's cpu_baseline computes the same function"""
    torch.manual_seed(3)
    q, k, v = torch.randn(1, 70, 4, 128), torch.randn(1, 90, 2, 128), torch.randn(1, 90, 2, 128)
    for causal in (False, True):
        o1, l1 = oracle_mod.attention_fwd(q.numpy(), k.numpy(), v.numpy(), causal)
        o2, l2 = oracle_mod.attention_fwd(q.numpy(), k.numpy(), v.numpy(), causal, fast=True)
        np.testing.assert_allclose(o1, o2, atol=1e-5)
        np.testing.assert_allclose(l1, l2, atol=1e-5)


def test_oracle_matches_torch_sdpa_config1(oracle_mod):
    """BASELINE config 1: b1 s512 h4 d128 fp32 through torch SDPA's CPU math path"""
    from torch.nn.attention import SDPBackend, sdpa_kernel
    torch.manual_seed(0)
    q, k, v = (torch.randn(1, 512, 4, 128) for _ in range(3))
    with sdpa_kernel(SDPBackend.MATH):
        ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    o, _ = oracle_mod.attention_fwd(q.numpy(), k.numpy(), v.numpy(), False)
    np.testing.assert_allclose(o, ref.transpose(1, 2).numpy(), atol=2e-5)
