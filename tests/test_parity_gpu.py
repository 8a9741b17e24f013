"""GPU parity tests (run with -m gpu on the B200 box).  Everything here calls the kernels through the C ABI
(include/fa_b200.h via tests/cabi.py) and checks them against
  * the committed golden vectors generated from the reference's own Python reference functions,
  * the CPU oracle (oracle/attn_oracle.c) on seeded inputs,
  * an fp32 torch reference on the GPU over the reference's shape grid (test_flash_attn.py:261-344),
  * size-independent properties at BASELINE.json's full sizes.
Tolerances: fp16 = the reference's own gates (test_flash_attn.py:407-414); bf16 = 8x those (see gpu_ref.GATES)."""
import os

import numpy as np
import pytest
import torch

import cabi
from conftest import golden_files, load_golden
from gpu_ref import GATES, assert_close, attention_ref, error_metrics

pytestmark = pytest.mark.gpu
DT = {"fp16": torch.float16, "bf16": torch.bfloat16}


def dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(dtype)


@pytest.mark.parametrize("path", golden_files("dense"), ids=lambda p: os.path.basename(p)[:-4])
def test_golden_dense(path, oracle_mod):
    g = load_golden(path)
    dt = DT[g["dtype"]]
    q, k, v, do = (dev(g[x], dt) for x in ("q", "k", "v", "dout"))
    o, lse = cabi.fwd(q, k, v, g["causal"])
    assert_close(o, dev(g["out"], torch.float32), dt, "O")
    o_ref, lse_ref = oracle_mod.attention_fwd(g["q"], g["k"], g["v"], g["causal"])
    assert (lse.cpu() - torch.from_numpy(lse_ref)).abs().max().item() <= 1e-3, "LSE"
    if "dq" in g:
        dq, dk, dv = cabi.bwd(q, k, v, o, lse, do, g["causal"])
        for name, x in (("dq", dq), ("dk", dk), ("dv", dv)):
            assert_close(x, dev(g[name], torch.float32), dt, name)


@pytest.mark.parametrize("path", golden_files("varlen"), ids=lambda p: os.path.basename(p)[:-4])
def test_golden_varlen(path):
    g = load_golden(path)
    dt = DT[g["dtype"]]
    q, k, v, do = (dev(g[x], dt) for x in ("q", "k", "v", "dout"))
    cu_q, cu_k = torch.from_numpy(g["cu_q"]).cuda(), torch.from_numpy(g["cu_k"]).cuda()
    kw = dict(cu_q=cu_q, cu_k=cu_k, max_sq=g["max_sq"], max_sk=g["max_sk"])
    o, lse = cabi.fwd(q, k, v, g["causal"], **kw)
    assert_close(o, dev(g["out"], torch.float32), dt, "O")
    dq, dk, dv = cabi.bwd(q, k, v, o, lse, do, g["causal"], **kw)
    for name, x in (("dq", dq), ("dk", dk), ("dv", dv)):
        assert_close(x, dev(g[name], torch.float32), dt, name)
    lens = np.diff(g["cu_q"])
    for b, n in enumerate(lens):  # LSE padding stays zero (flash_api.cpp:352)
        assert torch.all(lse[b, :, n:] == 0)


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("b,h,hk,sq,sk,d", [
    (1, 2, 1, 97, 203, 128), (2, 4, 2, 256, 256, 128), (1, 6, 3, 300, 77, 64), (1, 2, 2, 513, 640, 128),
])
def test_against_cpu_oracle_seeded(oracle_mod, b, h, hk, sq, sk, d, causal, dtype):
    dt = DT[dtype]
    torch.manual_seed(b * 7 + sq)
    q, k, v, do = (torch.randn(*s, device="cuda", dtype=dt) for s in ((b, sq, h, d), (b, sk, hk, d), (b, sk, hk, d), (b, sq, h, d)))
    o, lse = cabi.fwd(q, k, v, causal)
    qn, kn, vn, dn = (t.float().cpu().numpy() for t in (q, k, v, do))
    o_ref, lse_ref = oracle_mod.attention_fwd(qn, kn, vn, causal)
    assert_close(o, torch.from_numpy(o_ref).cuda(), dt, "O")
    assert (lse.cpu() - torch.from_numpy(lse_ref)).abs().max().item() <= 1e-3
    dq, dk, dv = cabi.bwd(q, k, v, o, lse, do, causal)
    r = oracle_mod.attention_bwd(qn, kn, vn, o_ref, lse_ref, dn, causal)
    for name, x, y in zip(("dq", "dk", "dv"), (dq, dk, dv), r):
        assert_close(x, torch.from_numpy(y).cuda(), dt, name)


# the reference's (seqlen_q, seqlen_k) grid, de-duplicated (test_flash_attn.py:261-344)
REF_SEQ = [(1, 1), (1, 63), (1, 64), (1, 65), (1, 127), (1, 128), (1, 129), (1, 1023), (1, 1024), (1, 1025),
           (63, 1), (64, 1), (65, 1), (127, 1), (128, 1), (129, 1), (1023, 1), (1024, 1), (1025, 1),
           (63, 63), (64, 64), (65, 65), (127, 127), (128, 128), (129, 129), (1023, 1023), (1024, 1024), (1025, 1025),
           (63, 64), (64, 63), (64, 65), (65, 64), (127, 128), (128, 127), (128, 129), (129, 128),
           (1023, 1024), (1024, 1023), (1024, 1025), (1025, 1024), (63, 128), (128, 63), (65, 127), (127, 65),
           (128, 1024), (1024, 128), (129, 1023), (1023, 129), (255, 257), (257, 255), (256, 512), (512, 256),
           (511, 513), (513, 511), (383, 385), (385, 383), (2, 3)]


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("nheads,nheads_k", [(2, 1), (6, 3), (4, 4)])
def test_reference_shape_grid(d, causal, nheads, nheads_k, dtype):
    """the reference's dense acceptance grid (GQA/MQA, ragged lengths) against the fp32 formulation its own tests use,
    in fp16 (the reference's dtype) and bf16 (the dtype of every BASELINE config); batch 3 for the small shapes, 1 for
    the 1k ones"""
    dt = DT[dtype]
    for i, (sq, sk) in enumerate(REF_SEQ):
        b = 3 if max(sq, sk) <= 257 else 1
        torch.manual_seed(1000 * i + d)
        q = torch.randn(b, sq, nheads, d, device="cuda", dtype=dt)
        k = torch.randn(b, sk, nheads_k, d, device="cuda", dtype=dt)
        v = torch.randn(b, sk, nheads_k, d, device="cuda", dtype=dt)
        do = torch.randn(b, sq, nheads, d, device="cuda", dtype=dt)
        o, lse = cabi.fwd(q, k, v, causal)
        ref = attention_ref(q, k, v, causal, do)
        tag = f"sq{sq} sk{sk}"
        assert_close(o, ref[0], dt, f"O {tag}")
        assert (lse - ref[1]).abs().max().item() <= 1e-3, f"LSE {tag}"
        dq, dk, dv = cabi.bwd(q, k, v, o, lse, do, causal)
        assert_close(dq, ref[2], dt, f"dq {tag}")
        assert_close(dk, ref[3], dt, f"dk {tag}")
        assert_close(dv, ref[4], dt, f"dv {tag}")


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_varlen_random_lengths(d, causal, dtype):
    """random per-sequence lengths as in the reference's varlen test (test_flash_attn.py:683-712); each sequence is
    compared with the dense reference evaluated on that sequence alone"""
    dt = DT[dtype]
    gen = torch.Generator().manual_seed(d + int(causal))
    for max_sq, max_sk, b, h, hk in [(129, 257, 4, 4, 2), (512, 300, 3, 6, 1), (64, 64, 5, 2, 2)]:
        lq = torch.randint(1, max_sq + 1, (b,), generator=gen)
        lk = torch.randint(1, max_sk + 1, (b,), generator=gen)
        lq[0], lk[-1] = max_sq, max_sk
        cu_q = torch.cat([torch.zeros(1, dtype=torch.int64), lq.cumsum(0)]).to(torch.int32).cuda()
        cu_k = torch.cat([torch.zeros(1, dtype=torch.int64), lk.cumsum(0)]).to(torch.int32).cuda()
        tq, tk = int(lq.sum()), int(lk.sum())
        torch.manual_seed(tq + tk)
        q = torch.randn(tq, h, d, device="cuda", dtype=dt)
        k = torch.randn(tk, hk, d, device="cuda", dtype=dt)
        v = torch.randn(tk, hk, d, device="cuda", dtype=dt)
        do = torch.randn(tq, h, d, device="cuda", dtype=dt)
        kw = dict(cu_q=cu_q, cu_k=cu_k, max_sq=max_sq, max_sk=max_sk)
        o, lse = cabi.fwd(q, k, v, causal, **kw)
        dq, dk, dv = cabi.bwd(q, k, v, o, lse, do, causal, **kw)
        for i in range(b):
            qs, qe, ks, ke = int(cu_q[i]), int(cu_q[i + 1]), int(cu_k[i]), int(cu_k[i + 1])
            ref = attention_ref(q[qs:qe][None], k[ks:ke][None], v[ks:ke][None], causal, do[qs:qe][None])
            assert_close(o[qs:qe][None], ref[0], dt, f"O seq{i}")
            assert (lse[i, :, : qe - qs] - ref[1][0]).abs().max().item() <= 1e-3
            assert_close(dq[qs:qe][None], ref[2], dt, f"dq seq{i}")
            assert_close(dk[ks:ke][None], ref[3], dt, f"dk seq{i}")
            assert_close(dv[ks:ke][None], ref[4], dt, f"dv seq{i}")


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties
# ---------------------------------------------------------------------------------------------------------------
FULL = [("C2", 4, 4096, False), ("C3", 4, 8192, True)]


@pytest.mark.parametrize("name,b,s,causal", FULL, ids=[f[0] for f in FULL])
def test_full_size_properties(name, b, s, causal):
    h, d, dt = 32, 128, torch.bfloat16
    torch.manual_seed(0)
    q = torch.randn(b, s, h, d, device="cuda", dtype=dt)
    k = torch.randn(b, s, h, d, device="cuda", dtype=dt)
    v = torch.randn(b, s, h, d, device="cuda", dtype=dt)
    o, lse = cabi.fwd(q, k, v, causal)
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all()

    # (1) every (batch, head) problem is independent: a slice run on its own is BIT-identical
    for (bi, hi) in [(0, 0), (b - 1, h - 1), (1, 17)]:
        qs, ks, vs = (t[bi:bi + 1, :, hi:hi + 1].contiguous() for t in (q, k, v))
        o1, l1 = cabi.fwd(qs, ks, vs, causal)
        assert torch.equal(o1, o[bi:bi + 1, :, hi:hi + 1]), f"slice ({bi},{hi}) differs from the batched run"
        assert torch.equal(l1, lse[bi:bi + 1, hi:hi + 1])
        # (2) the slice against the fp32 formulation, and no worse than 2x torch's own fused bf16 kernel
        ref_o, ref_l = attention_ref(qs, ks, vs, causal)
        assert_close(o1, ref_o, dt, f"O slice ({bi},{hi})")
        assert (l1 - ref_l).abs().max().item() <= 2e-3
        sd = torch.nn.functional.scaled_dot_product_attention(qs.transpose(1, 2), ks.transpose(1, 2), vs.transpose(1, 2),
                                                               is_causal=causal).transpose(1, 2)
        e_ours, e_torch = error_metrics(o1, ref_o), error_metrics(sd, ref_o)
        assert e_ours["mean_abs"] <= 2 * e_torch["mean_abs"] + 1e-5, (e_ours, e_torch)

    # (3) exact homogeneity in V: scaling V by a power of two scales O by exactly that factor
    o2, l2 = cabi.fwd(q, k, v * 2, causal)
    assert torch.equal(o2, o * 2) and torch.equal(l2, lse)

    # (4) each output row is a convex combination of value rows
    vmax = v.float().amax(dim=1, keepdim=True)
    vmin = v.float().amin(dim=1, keepdim=True)
    assert (o.float() <= vmax + 2e-2).all() and (o.float() >= vmin - 2e-2).all()

    # (5) non-causal: permuting the keys (with their values) does not change the result beyond accumulation order
    if not causal:
        perm = torch.randperm(s, device="cuda")
        o3, l3 = cabi.fwd(q[:1], k[:1, perm].contiguous(), v[:1, perm].contiguous(), False)
        assert (o3.float() - o[:1].float()).abs().max().item() <= 2e-2
        assert (l3 - lse[:1]).abs().max().item() <= 1e-3
    else:
        # causal: row i must not depend on keys j > i — corrupt the last keys and compare the early rows bit-exactly
        k2, v2 = k[:1].clone(), v[:1].clone()
        k2[:, s // 2:] = 7.0
        v2[:, s // 2:] = -3.0
        o4, l4 = cabi.fwd(q[:1], k2, v2, True)
        assert torch.equal(o4[:, : s // 2], o[:1, : s // 2]) and torch.equal(l4[:, :, : s // 2], lse[:1, :, : s // 2])


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_rescale_path_scores_growing_along_the_keys(d, causal, dtype):
    """The forward keeps a lazily updated reference max (it only moves when a key tile's exponentials would leave the
    2^9 range).  Random inputs almost never move it after the first tile, so this case makes every 128-key tile ~20
    exponent units larger than the one before: the rescale path (true max, exchange between the two threads of a row,
    O / l rescale, exponentials redone) runs at every step, and one row in this draw jumps by 2^141 inside one tile
    (its speculative exponentials overflow).  That row exposed a real bug on the B200: the polynomial exp2 wrapped
    around for arguments > 128 and the tile-sum vote did not trip, so a whole key tile was dropped (O off by ~3, LSE
    by 68.7; profiles/r01s2_rescale_bug.log).  The inputs are ill-conditioned (softmax nearly one-hot, |k| up to ~170),
    so the gates here are absolute and wide — measured errors on the fixed kernel: bf16 O max 1.8e-2 / mean 4e-3,
    LSE 1e-4 — but a dropped or mis-scaled tile is an O(1) error."""
    dt = DT[dtype]
    torch.manual_seed(7)
    b, sq, sk, h, hk = 2, 300, 1024, 4, 2
    q = torch.randn(b, sq, h, d, device="cuda", dtype=dt)
    k = torch.randn(b, sk, hk, d, device="cuda", dtype=dt)
    v = torch.randn(b, sk, hk, d, device="cuda", dtype=dt)
    grow = (1.0 + 6.0 * (torch.arange(sk, device="cuda") // 128)).to(dt)      # 1, 7, 13, ... per key tile
    k = (k * grow[None, :, None, None]).contiguous()

    def check(kk, what):
        o, lse = cabi.fwd(q, kk, v, causal)
        ro, rl = attention_ref(q, kk, v, causal)
        assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all(), what
        m = error_metrics(o, ro)
        assert m["max_abs"] <= 0.25 * max(1.0, m["ref_max"] / 4.0), f"{what}: O {m}"
        assert m["mean_abs"] <= 2e-2, f"{what}: O {m}"
        lerr = (lse - rl).abs().max().item()
        assert lerr <= 0.05 * max(1.0, rl.abs().max().item() / 8.0), f"{what}: LSE max err {lerr}"

    check(k, "growing")
    check(torch.flip(k, dims=[1]).contiguous(), "shrinking")   # reference set by the first tile, never moves again


def test_full_size_backward_c4_slice():
    """C4 is b4 s16384 fwd+bwd; the backward is checked on a (batch, head) slice at s=2048 against fp32 autograd.
    dK and dV must be bit-identical between two runs.  dQ of the default (fused) backward is accumulated with fp32
    reductions whose order is not fixed, so two runs may differ in the last bit; with no workspace the C ABI runs the two
    deterministic kernels (the reference's structure, flash_bwd_kernel.h) and dQ must be bit-identical as well."""
    dt = torch.bfloat16
    torch.manual_seed(4)
    q, k, v, do = (torch.randn(1, 2048, 2, 128, device="cuda", dtype=dt) for _ in range(4))
    for causal in (False, True):
        o, lse = cabi.fwd(q, k, v, causal)
        g1 = cabi.bwd(q, k, v, o, lse, do, causal)
        g2 = cabi.bwd(q, k, v, o, lse, do, causal)
        ref = attention_ref(q, k, v, causal, do)
        for name, a, b_, r in zip(("dq", "dk", "dv"), g1, g2, ref[2:]):
            if name != "dq":
                assert torch.equal(a, b_), f"{name} not deterministic"
            else:
                assert (a.float() - b_.float()).abs().max().item() <= 2.0 ** -7 * r.abs().max().item(), "dq run-to-run"
            assert_close(a, r, dt, name)
            assert_close(b_, r, dt, name)
        d1 = cabi.bwd(q, k, v, o, lse, do, causal, use_workspace=False)
        d2 = cabi.bwd(q, k, v, o, lse, do, causal, use_workspace=False)
        for name, a, b_, r in zip(("dq", "dk", "dv"), d1, d2, ref[2:]):
            assert torch.equal(a, b_), f"{name} not deterministic on the two-kernel path"
            assert_close(a, r, dt, name)


def test_launch_count_and_no_fallback():
    lib = cabi.load()
    q = torch.randn(1, 256, 2, 128, device="cuda", dtype=torch.bfloat16)
    o, lse = cabi.fwd(q, q, q, False)
    assert lib.fa_b200_last_launch_count() == 1
    dq, dk, dv = cabi.bwd(q, q, q, o, lse, q, False)
    assert lib.fa_b200_last_launch_count() >= 3


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
@pytest.mark.parametrize("b,h,hk,sq,sk,causal", [
    (2, 16, 4, 330, 271, False), (3, 8, 2, 41, 5452, False), (4, 4, 4, 1368, 174, False), (3, 3, 3, 631, 1286, True),
    (1, 12, 3, 1187, 240, True), (36, 2, 1, 263, 493, False), (2, 8, 2, 2048, 2048, True), (1, 4, 4, 64, 64, False),
])
def test_fused_backward_head_dim_64(dtype, b, h, hk, sq, sk, causal):
    """head_dim 64 through the fused dQ/dK/dV kernel (dQ^T is an M = 64 accumulator on TMEM lanes 0-15 of every quadrant)
    against the two deterministic kernels (workspace = NULL, the reference's structure flash_bwd_kernel.h:28-838 / :842-1676)
    and the fp32 reference: ragged tails, GQA group sums, more keys than queries and the reverse, causal."""
    dt = DT[dtype]
    torch.manual_seed(64 + sq)
    q = torch.randn(b, sq, h, 64, device="cuda").to(dt)
    k = torch.randn(b, sk, hk, 64, device="cuda").to(dt)
    v = torch.randn(b, sk, hk, 64, device="cuda").to(dt)
    do = torch.randn_like(q)
    o, lse = cabi.fwd(q, k, v, causal)
    lib = cabi.load()
    p = cabi.make_fwd_params(q, k, v, o, lse, causal)
    import ctypes
    if lib.fa_b200_bwd_workspace_bytes(ctypes.byref(p)) == 0:
        pytest.skip("FA_B200_BWD / FA_B200_BWD_D64 select the two-kernel backward")
    fused = cabi.bwd(q, k, v, o, lse, do, causal)
    det = cabi.bwd(q, k, v, o, lse, do, causal, use_workspace=False)
    ref = attention_ref(q, k, v, causal, do)
    ulp = 2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10
    for name, a, c, r in zip(("dq", "dk", "dv"), fused, det, ref[2:]):
        assert_close(a, r, dt, name + " (fused)")
        assert_close(c, r, dt, name + " (two kernels)")
        assert (a.float() - c.float()).abs().max().item() <= 2 * ulp * max(1.0, r.abs().max().item()), f"{name}: fused vs two kernels"


@pytest.mark.parametrize("b,sq,sk,h,hk,causal", [(5, 2211, 1202, 8, 4, True), (6, 2070, 1777, 8, 4, False), (4, 2275, 1261, 16, 4, True)])
def test_retry_pass_with_several_items_per_cta(b, sq, sk, h, hk, causal):
    """head_dim 64, fp16, scores ~ N(0, 16^2): nearly every item overflows its speculative steps and every CTA holds 3-4
    items, so pass 1 (exact steps) runs over several items per CTA.  The first retry scheme of round 2 hung here (a retry
    list whose length changed while pass 1 was running, profiles/r02_run31_gdb_hang.log); run in a subprocess under a timeout."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "diag_fwd_hang.py"), "0", "0", str(b), str(sq), str(sk), str(h),
                        str(hk), "64", str(int(causal)), "fp16", "4.0", "--check"], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and "FWDDIAG ok" in r.stdout and "CHECK ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
