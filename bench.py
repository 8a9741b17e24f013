#!/usr/bin/env python
"""bench.py — headline benchmark: attention forward TFLOP/s at head_dim=128 (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C4fwd|C5shard]

A "step" is one forward pass (one launch of flash_fwd_kernel_sm100_p4 through the C ABI) over one synthetic batch.
Default workload = BASELINE.json configs[1] ("C2": b4 s4096 h32 d128 bf16 forward, non-causal) on every rank; for
N > 1 the (batch x head) problems are sharded by batch — each rank owns an independent b=4 slab, no collective on
the data path — so scaling is "weak" and `value` = all ranks' FLOPs / max-over-ranks device time.

Printed JSON (one line, rank 0): the driver contract plus
  roofline     : tensor-bound; achieved = algorithmic FLOPs per launch / mean launch time (CUDA events on the
                 launching stream over the timed region); peak = MEASURED_PEAKS.json bf16_tflops (burst: the kernel is
                 timed alone in a ~tens-of-ms loop), "of measured"
  e2e          : same metric through the public API (flash_attn_turing.fwd_host) with HOST pinned buffers: H2D of q,k,v
                 and D2H of o,lse inside the timed region (chunked by batch, copies overlapped with the kernel)
  cpu_baseline : torch SDPA CPU math path (fp32) — north_star's named baseline — on a bounded slice of the same
                 workload, all host threads; the C oracle's float variant is timed beside it
`--impl reference` times that CPU arm alone (the reference ships no CPU implementation and no sm_100 build; its
kernels cannot run on the host — DESIGN.md §Measurement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
try:  # torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs must use every host core this process may run on
    _NCORES = len(os.sched_getaffinity(0))
except AttributeError:
    _NCORES = os.cpu_count() or 1
if int(os.environ.get("RANK", "0")) == 0:
    os.environ["OMP_NUM_THREADS"] = str(_NCORES)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-turing_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (batch per rank, seqlen, heads, head_dim, causal, description)
    "C2": (4, 4096, 32, 128, False, "bs=4 seq=4096 heads=32 hdim=128 bf16 forward"),
    "C3": (4, 8192, 32, 128, True, "bs=4 seq=8192 heads=32 hdim=128 bf16 causal forward"),
    "C4fwd": (4, 16384, 32, 128, False, "bs=4 seq=16384 heads=32 hdim=128 bf16 forward (fwd half of config 4)"),
    "C5shard": (32, 16384, 32, 128, False, "bs=256/8 seq=16384 heads=32 hdim=128 bf16 forward, one rank's batch shard"),
}
METRIC = "attention fwd TFLOP/s at head_dim=128; % of B200 bf16 tensor-core peak"


def fwd_flops(b, s, h, d, causal):
    return 4.0 * b * h * s * s * d * (0.5 if causal else 1.0)


def fwd_bytes(b, s, h, d):
    return 2 * (4 * b * s * h * d) + 4 * b * h * s  # q,k,v read + o written (16-bit) + lse fp32


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        return {"burst": float(m["bf16_tflops"]), "sustained": float(m.get("bf16_tflops_sustained", 0)),
                "hbm_gbs": float(m.get("hbm_gbs", 0)), "source": "MEASURED_PEAKS.json (of measured)"}
    except Exception:
        return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


def bind_near_gpu(index):
    """Restrict this process to the CPUs next to GPU `index` (NVML affinity mask) so that pinned host buffers are
    first-touched on the GPU's NUMA node; returns the previous affinity (None if nothing was changed).  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:   # CUDA_VISIBLE_DEVICES may reorder devices: look the GPU up by its PCI address
            import torch
            pr = torch.cuda.get_device_properties(index)
            hnd = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        except Exception:
            hnd = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
        near = {64 * w + bit for w, word in enumerate(words) for bit in range(64) if (word >> bit) & 1}
        prev = os.sched_getaffinity(0)
        allowed = near & set(prev)
        if allowed and allowed != set(prev):
            os.sched_setaffinity(0, allowed)
            return prev
    except Exception:
        pass
    return None


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_sdpa_baseline(s, d, causal, budget_s=12.0):
    """torch SDPA CPU math path, fp32, on a bounded slice (b=1, a few heads) of the workload; all host threads"""
    import torch
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel
    heads = 4
    torch.set_num_threads(_NCORES)
    torch.manual_seed(0)
    q, k, v = (torch.randn(1, heads, s, d) for _ in range(3))
    fl = fwd_flops(1, s, heads, d, causal)
    with sdpa_kernel(SDPBackend.MATH), torch.no_grad():
        F.scaled_dot_product_attention(q, k, v, is_causal=causal)  # warm-up
        t0 = time.perf_counter(); n = 0
        while True:
            F.scaled_dot_product_attention(q, k, v, is_causal=causal)
            n += 1
            dt = time.perf_counter() - t0
            if dt > budget_s or n >= 50:
                break
    tf = fl * n / dt / 1e12
    # the C oracle's float32 variant on the same slice, for context
    oracle_tf = None
    try:
        from oracle import oracle
        qn, kn, vn = (t.permute(0, 2, 1, 3).contiguous().numpy() for t in (q, k, v))
        oracle.attention_fwd(qn[:, :256], kn, vn, causal, fast=True)
        t0 = time.perf_counter()
        oracle.attention_fwd(qn, kn, vn, causal, fast=True)
        oracle_tf = fl / (time.perf_counter() - t0) / 1e12
    except Exception:
        pass
    return {"value": tf, "unit": "TFLOP/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"torch SDPA CPU math path fp32, b=1 h={heads} s={s} d={d} causal={causal} slice of the workload, "
                      f"{n} iters in {dt:.1f}s; C oracle (oracle_fwd_f32, OpenMP) on the same slice: "
                      + (f"{oracle_tf:.4f} TFLOP/s" if oracle_tf else "n/a")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=list(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--bwd", action="store_true", help="also time the backward (dot + dQ + dK/dV kernels) and report it under \"bwd\"")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    b, s, h, d, causal, desc = CONFIGS[args.config]

    if args.impl == "reference":
        if rank != 0:
            return 0
        base = cpu_sdpa_baseline(s, d, causal, budget_s=max(5.0, min(60.0, 2.0 * args.steps)))
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.config}: {desc}", "note": "CPU arm: the reference has no CPU path and no sm_100 "
                           "build; north_star names torch SDPA's CPU math path as the host baseline"},
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    import cabi
    import flash_attn_turing as fat

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    dt = torch.bfloat16
    torch.manual_seed(1000 + rank if world > 1 else 0)
    q = torch.randn(b, s, h, d, device=dev, dtype=dt)
    k = torch.randn(b, s, h, d, device=dev, dtype=dt)
    v = torch.randn(b, s, h, d, device=dev, dtype=dt)
    o = torch.empty_like(q)
    lse = torch.empty(b, h, s, device=dev, dtype=torch.float32)
    lib = cabi.load()
    import ctypes
    params = cabi.make_fwd_params(q, k, v, o, lse, causal)
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)

    def step():
        rc = lib.fa_b200_fwd(ctypes.byref(params), sptr)
        if rc != 0:
            raise RuntimeError(lib.fa_b200_last_error().decode())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    launches_per_step = lib.fa_b200_last_launch_count()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    # inputs (q,k,v,o = 512 MiB at C2) exceed the 126 MB L2, so every step streams them from HBM again
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    sync_all()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    flops_rank = fwd_flops(b, s, h, d, causal)
    value = flops_rank * world / (ms_per_step * 1e-3) / 1e12

    # ---- optional: backward leg (BASELINE config 4 is fwd+bwd); reported beside the headline, not inside it ----
    bwd = None
    if args.bwd:
        do = torch.randn_like(q)
        n_b = max(3, min(args.steps, 10))
        for _ in range(2):
            g = fat.bwd(q, k, v, o, lse, do, causal)
        bwd_launches = fat.last_launch_count()
        sync_all()
        e0.record(stream)
        for _ in range(n_b):
            fat.bwd(q, k, v, o, lse, do, causal)
        e1.record(stream)
        sync_all()
        tb = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        ms_b = float(tb.item()) / n_b
        bwd = {"ms_per_step": ms_b, "value": 2.5 * flops_rank * world / (ms_b * 1e-3) / 1e12, "unit": "TFLOP/s (algorithmic 2.5x fwd)",
               "steps": n_b, "gpu_launches_per_step": bwd_launches,
               "kernels": "flash_bwd_dot_do_o_kernel_sm100 + flash_bwd_dk_dv_kernel_sm100_fused (dQ through fp32 bulk reductions) + "
                          "flash_bwd_dq_kernel_sm100_convert; FA_B200_BWD=det selects the two deterministic kernels"}
        del do, g

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        prev_aff = bind_near_gpu(local_rank)   # host buffers on the GPU's NUMA node (restored below for the CPU baseline)
        hq, hk, hv = (torch.randn(b, s, h, d, dtype=dt).pin_memory() for _ in range(3))
        ho = torch.empty(b, s, h, d, dtype=dt).pin_memory()
        hl = torch.empty(b, h, s, dtype=torch.float32).pin_memory()
        host_fwd = fat.HostForward()   # batch-chunked H2D / kernel / D2H pipeline over three streams (hostio.py)

        def e2e_step():
            host_fwd(hq, hk, hv, causal, out=ho, lse=hl, sync=False)

        n_e2e = max(3, min(args.steps, 10))
        for _ in range(2):
            e2e_step()
        sync_all()
        e0.record(stream)
        for _ in range(n_e2e):
            e2e_step()
        e1.record(stream)
        sync_all()
        t2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_e2e = float(t2.item()) / n_e2e
        e2e = {"value": flops_rank * world / (ms_e2e * 1e-3) / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": 3 * q.numel() * 2, "d2h_bytes_per_step": o.numel() * 2 + lse.numel() * 4,
               "ms_per_step": ms_e2e, "steps": n_e2e, "api": "flash_attn_turing.fwd_host(q,k,v,is_causal) on pinned host buffers: per-batch chunks, H2D / kernel / "
                      "D2H overlapped on three streams; every step moves all inputs up and all outputs down",
               "gpu_launches_per_step": b, "numa_bound": prev_aff is not None}
        if prev_aff is not None:
            os.sched_setaffinity(0, prev_aff)

    if rank == 0:
        peaks = measured_peaks()
        achieved = flops_rank / (ms_total / args.steps * 1e-3) / 1e12  # this rank's kernel, mean launch time
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(args.config)
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.config}: {desc}", "batch_per_gpu": b, "seq_len": s, "heads": h, "head_dim": d,
                       "causal": causal, "parallelism": f"batch-shard x{world} (independent (batch,head) problems, no collective)",
                       "l2": "inputs+outputs 512 MiB+ per step exceed the 126 MB L2 (no flush needed)",
                       "algorithmic_bytes_per_step": fwd_bytes(b, s, h, d)},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["burst"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["burst"], "traffic": traffic, "peak_source": peaks["source"],
                         "frac_of_sustained": (achieved / peaks["sustained"]) if peaks["sustained"] else None,
                         "frac_of_nominal_2250": achieved / 2250.0, "kernel": "flash_fwd_kernel_sm100_p4<bf16> (FA_B200_FWD / FA_B200_EMU select the A/B variants)"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if bwd:
            line["bwd"] = bwd
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sdpa_baseline(s, d, causal)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
