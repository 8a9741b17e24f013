"""Generate tests/golden/*.npz from the REFERENCE's own Python reference functions.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference ships no golden vectors: its tests draw unseeded randn and compare the CUDA kernels
with `vanilla_attention_ref` / `memory_efficient_attention_ref` (/root/reference/test_flash_attn.py:134-248).  This
script imports that very file (with a stub standing in for the CUDA extension it imports at the top), feeds seeded
inputs through those two functions on the CPU in float32, and stores inputs + outputs.  tests/test_oracle.py pins
oracle/attn_oracle.c against these files; the GPU parity tests use them as fixtures too.

Inputs are drawn as randn and rounded to the storage dtype (fp16 or bf16) first, then held as float32 — exactly
the values a kernel sees.  Dense cases go through BOTH reference functions (they must agree); varlen cases follow
the reference's per-sequence loop (test_flash_attn.py:790-811).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TEST = "/root/reference/test_flash_attn.py"


def load_reference_module():
    stub = types.ModuleType("flash_attn_turing")
    for name in ("fwd", "bwd", "varlen_fwd", "varlen_bwd"):
        setattr(stub, name, None)
    saved = sys.modules.get("flash_attn_turing")
    sys.modules["flash_attn_turing"] = stub
    try:
        spec = importlib.util.spec_from_file_location("ref_test_flash_attn", REF_TEST)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            del sys.modules["flash_attn_turing"]
        else:
            sys.modules["flash_attn_turing"] = saved
    return mod


# (name, b, h, h_k, sq, sk, d, causal, dtype)   — shapes drawn from the reference's own grid (test_flash_attn.py:251-344)
DENSE = [
    ("d64_1x1", 1, 2, 1, 1, 1, 64, False, "fp16"),
    ("d128_1x1_c", 1, 2, 1, 1, 1, 128, True, "fp16"),
    ("d128_63x65_c", 1, 6, 3, 63, 65, 128, True, "fp16"),
    ("d64_65x63_c", 1, 6, 1, 65, 63, 64, True, "fp16"),       # sq > sk: leading rows see no key -> 0
    ("d128_128x128", 1, 4, 2, 128, 128, 128, False, "fp16"),
    ("d128_128x128_c", 1, 4, 2, 128, 128, 128, True, "bf16"),
    ("d128_129x127", 1, 2, 1, 129, 127, 128, False, "bf16"),
    ("d64_127x129_c", 2, 4, 2, 127, 129, 64, True, "bf16"),
    ("d128_257x300_c", 1, 2, 2, 257, 300, 128, True, "bf16"),
    ("d128_1x257", 1, 2, 2, 1, 257, 128, False, "fp16"),
    ("d128_300x1_c", 1, 2, 1, 300, 1, 128, True, "fp16"),
    ("d128_c1_512", 1, 4, 4, 512, 512, 128, False, "bf16"),    # BASELINE config 1 (b1 s512 h4 d128); out only
    ("d128_c1_512_c", 1, 2, 1, 512, 512, 128, True, "bf16"),
]
# (name, h, h_k, seqlens_q, seqlens_k, d, causal, dtype)
VARLEN = [
    ("v_d128_a", 2, 1, [3, 128, 65, 1], [130, 7, 64, 1], 128, False, "fp16"),
    ("v_d128_a_c", 4, 2, [3, 128, 65, 1], [130, 7, 64, 1], 128, True, "fp16"),
    ("v_d64_b_c", 6, 1, [200, 17], [31, 257], 64, True, "bf16"),
    ("v_d128_c", 2, 2, [129, 255, 64], [129, 255, 64], 128, True, "bf16"),
]


def store16(x, dtype):
    """inputs are stored as their 16-bit patterns (fp16 -> float16 array, bf16 -> uint16 bit pattern)"""
    if dtype == "fp16":
        return x.numpy().astype(np.float16)
    return x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def quantize(x, dtype):
    return x.to(torch.float16 if dtype == "fp16" else torch.bfloat16).to(torch.float32)


def main():
    ref = load_reference_module()
    torch.manual_seed(20261017)
    torch.set_num_threads(8)
    for name, b, h, hk, sq, sk, d, causal, dtype in DENSE:
        q = quantize(torch.randn(b, sq, h, d), dtype)
        k = quantize(torch.randn(b, sk, hk, d), dtype)
        v = quantize(torch.randn(b, sk, hk, d), dtype)
        do = quantize(torch.randn(b, sq, h, d), dtype)
        o1, dq1, dk1, dv1 = ref.vanilla_attention_ref(q, k, v, do, causal)
        o2, dq2, dk2, dv2 = ref.memory_efficient_attention_ref(q, k, v, do, causal)
        # the two reference formulations must agree wherever SDPA defines a value (rows with no visible key are
        # NaN in torch SDPA and 0 in vanilla_attention_ref / the kernels, test_flash_attn.py:171)
        ok = torch.isfinite(o2)
        assert torch.allclose(o1[ok], o2[ok], atol=2e-5, rtol=1e-4), name
        np.savez_compressed(
            os.path.join(HERE, f"dense_{name}.npz"),
            q=store16(q, dtype), k=store16(k, dtype), v=store16(v, dtype), dout=store16(do, dtype),
            out=o1.detach().numpy(), causal=np.array(causal), dtype=np.array(dtype),
            **({} if "c1_512" in name else dict(dq=dq1.numpy(), dk=dk1.numpy(), dv=dv1.numpy())))
        print("dense", name, "ok")
    for name, h, hk, sql, skl, d, causal, dtype in VARLEN:
        tq, tk = sum(sql), sum(skl)
        q = quantize(torch.randn(tq, h, d), dtype)
        k = quantize(torch.randn(tk, hk, d), dtype)
        v = quantize(torch.randn(tk, hk, d), dtype)
        do = quantize(torch.randn(tq, h, d), dtype)
        cu_q = np.concatenate([[0], np.cumsum(sql)]).astype(np.int32)
        cu_k = np.concatenate([[0], np.cumsum(skl)]).astype(np.int32)
        outs, dqs, dks, dvs = [], [], [], []
        for i in range(len(sql)):
            qi = q[cu_q[i]:cu_q[i + 1]].unsqueeze(0).contiguous()
            ki = k[cu_k[i]:cu_k[i + 1]].unsqueeze(0).contiguous()
            vi = v[cu_k[i]:cu_k[i + 1]].unsqueeze(0).contiguous()
            di = do[cu_q[i]:cu_q[i + 1]].unsqueeze(0).contiguous()
            o_i, dq_i, dk_i, dv_i = ref.vanilla_attention_ref(qi, ki, vi, di, causal)
            outs.append(o_i.squeeze(0)); dqs.append(dq_i.squeeze(0)); dks.append(dk_i.squeeze(0)); dvs.append(dv_i.squeeze(0))
        np.savez_compressed(
            os.path.join(HERE, f"varlen_{name}.npz"),
            q=store16(q, dtype), k=store16(k, dtype), v=store16(v, dtype), dout=store16(do, dtype),
            cu_q=cu_q, cu_k=cu_k, max_sq=np.array(max(sql)), max_sk=np.array(max(skl)),
            out=torch.cat(outs).detach().numpy(), dq=torch.cat(dqs).numpy(), dk=torch.cat(dks).numpy(),
            dv=torch.cat(dvs).numpy(), causal=np.array(causal), dtype=np.array(dtype))
        print("varlen", name, "ok")


if __name__ == "__main__":
    main()
