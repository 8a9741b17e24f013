#!/usr/bin/env python
"""bench.py — headline benchmark: attention forward TFLOP/s at head_dim=128 (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C4fwd|C5shard] [--no-legs]

A "step" is one forward pass (one launch of flash_fwd_kernel_sm100_p4 through the C ABI, include/fa_b200.h) over one
synthetic batch resident in HBM.

Headline workload (`value`, `config.workload`):
  N = 1 : BASELINE.json configs[1] ("C2": b4 s4096 h32 d128 bf16 forward, non-causal) — the configuration `metric` is
          quoted on.  The same line also carries, each timed in its own loop with its own clock samples:
            sustained : C2 again in a >= 2 s loop (the 1 kW power cap governs; fraction of the sustained peak)
            configs   : C3 (b4 s8192 causal), C4fwd, C4bwd (b4 s16384 forward / backward), D64fwd, D64bwd (b4 s4096 head_dim 64), C5shard (b32 s16384: one
                        rank's slab of config 5)
  N > 1 : the same C2 workload on every rank (b=4 per rank, seed 1000 + rank; batch x head problems shard with no
          collective on the data path), so scaling is "weak" in the contract's sense — per-GPU work is the N=1 work — and
          `value` = all ranks' FLOPs / max-over-ranks device time.  The line also carries
            configs.C5shard : BASELINE.json configs[4] ("C5": b256 s16384 h32 d128 over 8 GPUs) — every rank runs one
                        b=32 s=16384 slab (2^31 elements per tensor, ~117 ms per step, power-capped clocks), timed
                        between barriers, max over ranks, whole-job TFLOP/s; the N=1 line's configs.C5shard is the
                        single-GPU point of the same per-rank workload, so config 5 has values at N = 1, 2, 4, 8
            shard_io  : NCCL scatter of Q/K/V + gather of O/LSE through flash_attn_turing.sharded (outside the headline)

Printed JSON (one line, rank 0): the driver contract plus
  roofline     : tensor-bound; achieved = algorithmic FLOPs per launch / mean launch time (CUDA events on the
                 launching stream over the timed region); peak = MEASURED_PEAKS.json bf16_tflops (burst) for loops of
                 tens of ms, bf16_tflops_sustained beside it — "of measured"
  e2e          : the C2 metric through the public API (flash_attn_turing.fwd_host) with HOST pinned buffers: H2D of q,k,v
                 and D2H of o,lse inside the timed region (chunked, copies overlapped with the kernel); copy_floor_ms is
                 the same traffic with no kernel
  cpu_baseline : torch SDPA CPU math path (fp32) — north_star's named baseline — on a bounded slice of the same
                 workload, all host threads; the C oracle's float variant is timed beside it
`--impl reference` times that CPU arm alone (the reference ships no CPU implementation and no sm_100 build; its
kernels cannot run on the host — DESIGN.md §Measurement).
"""
import argparse
import ctypes
import importlib.util
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
PKG_ROOT = os.path.join(ROOT, "flash-attention-turing_b200")
sys.path.insert(0, PKG_ROOT)
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (batch per rank, seqlen, heads, head_dim, causal, description)
    "C2": (4, 4096, 32, 128, False, "bs=4 seq=4096 heads=32 hdim=128 bf16 forward"),
    "C3": (4, 8192, 32, 128, True, "bs=4 seq=8192 heads=32 hdim=128 bf16 causal forward"),
    "C4fwd": (4, 16384, 32, 128, False, "bs=4 seq=16384 heads=32 hdim=128 bf16 forward (fwd half of config 4)"),
    "C5shard": (32, 16384, 32, 128, False, "bs=256/8 seq=16384 heads=32 hdim=128 bf16 forward, one rank's batch shard of config 5"),
    "D64": (4, 4096, 32, 64, False, "bs=4 seq=4096 heads=32 hdim=64 bf16 forward (the reference's other head_dim, benchmark.sh:20)"),
}
METRIC = "attention fwd TFLOP/s at head_dim=128; % of B200 bf16 tensor-core peak"
CPU_SLICE_HEADS = 4


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


def load_cabi():
    """the ctypes binding of include/fa_b200.h (flash_attn_turing/cabi.py), loaded by path"""
    spec = importlib.util.spec_from_file_location("fa_b200_cabi", os.path.join(PKG_ROOT, "flash_attn_turing", "cabi.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


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
    """samples nvidia-smi clocks / throttle reasons while a timed region runs"""
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
            time.sleep(0.1)
        except Exception:
            self.proc = None
        return self

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
        # median over the samples taken under load (power above the idle floor), else over all
        load = [c for c, w in zip(sm, pw) if w > 0.6 * max(pw)] or sm
        load.sort()
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_sdpa_baseline(s, d, causal, budget_s=12.0):
    """torch SDPA CPU math path, fp32, on a bounded slice (b=1, a few heads) of the workload; all host threads"""
    import torch
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel
    heads = CPU_SLICE_HEADS
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
            "sample": f"torch SDPA CPU math path fp32 on a b=1 h={heads} slice (s={s} d={d} causal={causal}) of the workload, rate "
                      f"compared, {n} iters in {dt:.1f}s; C oracle (oracle_fwd_f32, OpenMP) on the same slice: "
                      + (f"{oracle_tf:.4f} TFLOP/s" if oracle_tf else "n/a")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=list(CONFIGS), help="headline workload (default: C2, per rank)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the sustained / other-config legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg_name = args.config or "C2"
    b, s, h, d, causal, desc = CONFIGS[cfg_name]

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the CPU arm is quoted on the N=1 headline workload (C2) whatever N is: it runs on rank 0's host cores only
        rb, rs, rh, rd, rcausal, rdesc = CONFIGS[args.config or "C2"]
        base = cpu_sdpa_baseline(rs, rd, rcausal, budget_s=max(5.0, min(60.0, 2.0 * args.steps)))
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"b=1 h={CPU_SLICE_HEADS} slice of {args.config or 'C2'} ({rdesc}): rate on the slice, "
                                       "CPU arm on rank 0's host cores only (does not scale with --gpus)",
                           "slice": {"batch": 1, "heads": CPU_SLICE_HEADS, "seq_len": rs, "head_dim": rd, "causal": rcausal},
                           "note": "the reference has no CPU path and no sm_100 build; north_star names torch SDPA's CPU math "
                                   "path as the host baseline"},
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    cabi = load_cabi()
    import flash_attn_turing as fat

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    dt = torch.bfloat16
    lib = cabi.load()
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)
    peaks = measured_peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_inputs(cb, cs, ch, cd):
        g = torch.Generator(device=dev).manual_seed(1000 + rank if world > 1 else 0)
        return tuple(torch.empty(cb, cs, ch, cd, device=dev, dtype=dt).normal_(generator=g) for _ in range(3))

    def time_loop(fn, steps, warmup, min_seconds=0.0):
        """W warm-ups, then `steps` timed calls between two barriers (CUDA events on the launching stream), max over ranks;
        with min_seconds the loop length is first calibrated so that the timed region lasts at least that long"""
        for _ in range(warmup):
            fn()
        sync_all()
        if min_seconds > 0:
            e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
            steps = max(steps, int(min_seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)) + 1)
        sampler = ClockSampler(local_rank).start() if rank == 0 else None
        sync_all()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        sync_all()
        ms_local = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        return max_over_ranks(ms_local) / steps, ms_local / steps, steps, clocks

    def fwd_leg(name, steps, warmup, min_seconds=0.0, tensors=None):
        cb, cs, ch, cd, ccausal, cdesc = CONFIGS[name]
        q, k, v = tensors if tensors is not None else make_inputs(cb, cs, ch, cd)
        o = torch.empty_like(q)
        lse = torch.empty(cb, ch, cs, device=dev, dtype=torch.float32)
        params = cabi.make_fwd_params(q, k, v, o, lse, ccausal)

        def step():
            rc = lib.fa_b200_fwd(ctypes.byref(params), sptr)
            if rc != 0:
                raise RuntimeError(lib.fa_b200_last_error().decode())

        step()
        launches = lib.fa_b200_last_launch_count()
        ms, ms_local, n, clocks = time_loop(step, steps, warmup, min_seconds)
        fl = fwd_flops(cb, cs, ch, cd, ccausal)
        tf_rank = fl / (ms_local * 1e-3) / 1e12
        leg = {"workload": f"{name}: {cdesc}", "value": fl * world / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms,
               "steps": n, "warmup": warmup, "achieved_this_rank": tf_rank, "frac_of_burst_peak": tf_rank / peaks["burst"],
               "frac_of_sustained_peak": (tf_rank / peaks["sustained"]) if peaks["sustained"] else None,
               "timed_region_ms": ms * n, "gpu_launches": launches * n, "clocks": clocks,
               "algorithmic_bytes_per_step": fwd_bytes(cb, cs, ch, cd)}
        return leg, (q, k, v, o, lse)

    def bwd_leg(name, tensors, steps, desc):
        """backward of CONFIGS[name] through the C ABI on the forward leg's tensors: dot(dO,O) + fused dK/dV/dQ + dQ convert"""
        cb, cs, ch, cd, ccausal, _ = CONFIGS[name]
        qb, kb, vb, ob, lb = tensors
        dob = torch.randn_like(qb)
        dqb, dkb, dvb, dsb = torch.empty_like(qb), torch.empty_like(kb), torch.empty_like(vb), torch.empty_like(lb)
        bp = cabi.BwdParams()
        bp.fwd = cabi.make_fwd_params(qb, kb, vb, ob, lb, ccausal)
        bp.dout, bp.dq, bp.dk, bp.dv, bp.dsum = dob.data_ptr(), dqb.data_ptr(), dkb.data_ptr(), dvb.data_ptr(), dsb.data_ptr()
        nbytes = int(lib.fa_b200_bwd_workspace_bytes(ctypes.byref(bp.fwd)))
        ws = torch.empty(max(nbytes, 1), device=dev, dtype=torch.uint8)
        bp.workspace = ws.data_ptr() if nbytes > 0 else None

        def bstep():
            rc = lib.fa_b200_bwd(ctypes.byref(bp), sptr)
            if rc != 0:
                raise RuntimeError(lib.fa_b200_last_error().decode())

        bstep()
        bl = lib.fa_b200_last_launch_count()
        msb, _, nb, clk = time_loop(bstep, steps, 2)
        flb = 2.5 * fwd_flops(cb, cs, ch, cd, ccausal)
        tfb = flb / (msb * 1e-3) / 1e12
        return {"workload": desc, "value": tfb, "unit": "TFLOP/s (algorithmic 2.5x fwd)", "ms_per_step": msb, "steps": nb,
                "frac_of_burst_peak": tfb / peaks["burst"],
                "frac_of_sustained_peak": tfb / peaks["sustained"] if peaks["sustained"] else None,
                "timed_region_ms": msb * nb, "gpu_launches": bl * nb, "clocks": clk}

    # ---------------- headline ----------------
    head, (q, k, v, o, lse) = fwd_leg(cfg_name, args.steps, args.warmup)
    ms_per_step, value = head["ms_per_step"], head["value"]
    achieved = head["achieved_this_rank"]
    flops_rank = fwd_flops(b, s, h, d, causal)

    legs = {}
    sustained = None
    shard_io = None
    if not args.no_legs:
        if world == 1:
            # C2 again in a >= 2 s loop: the power cap, not the burst clock, governs (denominator: sustained cuBLAS peak)
            sustained, _ = fwd_leg(cfg_name, args.steps, 3, min_seconds=2.0, tensors=(q, k, v))
            del o, lse
            for name, st in (("C3", 20), ("C4fwd", 10)):
                if name == cfg_name:
                    continue
                if name == "C4fwd":
                    leg, (q4, k4, v4, o4, l4) = fwd_leg(name, st, 3)
                    legs[name] = leg
                    # config 4's backward half: dot(dO,O) + fused dK/dV/dQ kernel + dQ convert, through the C ABI
                    legs["C4bwd"] = bwd_leg("C4fwd", (q4, k4, v4, o4, l4), 5,
                                            "C4 backward: bs=4 seq=16384 heads=32 hdim=128 bf16 (dO.O + fused dK/dV/dQ + dQ convert)")
                    del q4, k4, v4, o4, l4
                else:
                    legs[name], _ = fwd_leg(name, st, 3)
                torch.cuda.empty_cache()
            if cfg_name != "D64":      # head_dim 64: forward and backward (fused dK/dV/dQ kernel with an M = 64 dQ^T)
                legs["D64fwd"], t64 = fwd_leg("D64", 20, 3)
                legs["D64bwd"] = bwd_leg("D64", t64, 10, "D64 backward: bs=4 seq=4096 heads=32 hdim=64 bf16 (dO.O + fused dK/dV/dQ + dQ convert)")
                del t64
                torch.cuda.empty_cache()
            if cfg_name != "C5shard":
                try:
                    legs["C5shard"], _ = fwd_leg("C5shard", 10, 2)   # the single-GPU point of the N>1 headline workload
                except torch.OutOfMemoryError as e:  # pragma: no cover
                    legs["C5shard"] = {"error": str(e)[:200]}
                torch.cuda.empty_cache()
        else:
            del q, k, v, o, lse
            torch.cuda.empty_cache()
            if cfg_name != "C5shard":
                # north_star's multi-GPU configuration: one b=32 s=16384 slab of config 5 per rank (power-capped regime)
                legs["C5shard"], _ = fwd_leg("C5shard", 10, 2)
                legs["C5shard"]["note"] = ("config 5 (b256 s16384 over 8 GPUs): per-rank slab b=32; value = all ranks' FLOPs / "
                                           "max-over-ranks time; compare with configs.C5shard of the N=1 line")
                torch.cuda.empty_cache()
            # shard I/O: rank 0 holds a whole b = 4*world batch of C2 rows, NCCL scatters Q/K/V and gathers O/LSE
            from flash_attn_turing import sharded
            bb = 4 * world
            shapes = ((bb, 4096, 32, 128), (bb, 4096, 32, 128))
            fq = fk = fv = None
            if rank == 0:
                fq, fk, fv = (torch.randn(*shapes[0], device=dev, dtype=dt) for _ in range(3))
            sharded.fwd_sharded(fq, fk, fv, False, fat.fwd, shapes=shapes, dtype=dt, device=dev)   # warm-up (NCCL channels)
            sync_all()
            t0 = time.perf_counter()
            e0.record(stream)
            oo, ll = sharded.fwd_sharded(fq, fk, fv, False, fat.fwd, shapes=shapes, dtype=dt, device=dev)
            e1.record(stream)
            sync_all()
            io_ms = max_over_ranks(e0.elapsed_time(e1))
            moved = (3 * bb * 4096 * 32 * 128 * 2 + bb * 4096 * 32 * 128 * 2 + bb * 32 * 4096 * 4) * (world - 1) / world
            shard_io = {"ms": io_ms, "wall_ms": (time.perf_counter() - t0) * 1e3, "bytes_over_nvlink": int(moved),
                        "gb_per_s": moved / (io_ms * 1e-3) / 1e9,
                        "what": f"flash_attn_turing.sharded.fwd_sharded on a b={bb} s4096 h32 d128 batch held by rank 0: NCCL send/recv "
                                "scatter of Q,K,V + per-rank forward + gather of O,LSE (includes the kernel, ~0.9 ms); outside the headline region"}
            del fq, fk, fv, oo, ll
            torch.cuda.empty_cache()

    # ---- end to end through the public API with host buffers (C2 shapes at every N) ----
    e2e = None
    if not args.no_e2e:
        eb, es, eh, ed, ecausal, _ = CONFIGS["C2"]
        prev_aff = bind_near_gpu(local_rank)   # host buffers on the GPU's NUMA node (restored below for the CPU baseline)
        hq, hk, hv = (torch.randn(eb, es, eh, ed, dtype=dt).pin_memory() for _ in range(3))
        ho = torch.empty(eb, es, eh, ed, dtype=dt).pin_memory()
        hl = torch.empty(eb, eh, es, dtype=torch.float32).pin_memory()
        host_fwd = fat.HostForward()   # chunked H2D / kernel / D2H pipeline over three streams (hostio.py)

        def e2e_step():
            host_fwd(hq, hk, hv, ecausal, out=ho, lse=hl, sync=False)

        def copy_only():
            host_fwd(hq, hk, hv, ecausal, out=ho, lse=hl, sync=False, copy_only=True)

        n_e2e = max(3, min(args.steps, 10))
        ms_e2e, _, _, _ = time_loop(e2e_step, n_e2e, 2)
        ms_floor, _, _, _ = time_loop(copy_only, n_e2e, 1)
        fl_e2e = fwd_flops(eb, es, eh, ed, ecausal)
        e2e = {"value": fl_e2e * world / (ms_e2e * 1e-3) / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": 3 * eb * es * eh * ed * 2, "d2h_bytes_per_step": eb * es * eh * ed * 2 + eb * eh * es * 4,
               "ms_per_step": ms_e2e, "steps": n_e2e, "copy_floor_ms": ms_floor, "frac_of_copy_floor": ms_floor / ms_e2e,
               "workload": "C2 (b4 s4096 h32 d128 bf16 forward) per rank at every N",
               "api": "flash_attn_turing.fwd_host(q,k,v,is_causal) on pinned host buffers: chunks of the batch, H2D / kernel / "
                      "D2H overlapped on three streams; every step moves all inputs up and all outputs down; copy_floor_ms = the "
                      "same copies with no kernel (what this host's PCIe / NUMA path can move)",
               "gpu_launches_per_step": host_fwd.last_chunks, "numa_bound": prev_aff is not None}
        if prev_aff is not None:
            os.sched_setaffinity(0, prev_aff)

    if rank == 0:
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(cfg_name)
        except Exception:
            pass
        long_loop = head["timed_region_ms"] > 500.0   # a timed region this long runs under the power cap
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{cfg_name}: {desc}" + ("" if world == 1 else f" x {world} ranks (b={b * world} in total)"),
                       "batch_per_gpu": b, "seq_len": s, "heads": h, "head_dim": d,
                       "causal": causal, "parallelism": f"batch-shard x{world} (independent (batch,head) problems, no collective)",
                       "l2": "inputs+outputs 512 MiB+ per step exceed the 126 MB L2 (no flush needed)",
                       "algorithmic_bytes_per_step": fwd_bytes(b, s, h, d)},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["sustained"] if long_loop else peaks["burst"],
                         "unit": "TFLOP/s", "frac": achieved / (peaks["sustained"] if long_loop else peaks["burst"]),
                         "peak_kind": "sustained (timed region > 0.5 s)" if long_loop else "burst (timed region of tens of ms)",
                         "traffic": traffic, "peak_source": peaks["source"],
                         "frac_of_burst": achieved / peaks["burst"],
                         "frac_of_sustained": (achieved / peaks["sustained"]) if peaks["sustained"] else None,
                         "frac_of_nominal_2250": achieved / 2250.0,
                         "kernel": "flash_fwd_kernel_sm100_p4<128, bf16> (one launch per step)"},
            "gpu_launches": head["gpu_launches"],
            "clocks": head["clocks"],
        }
        if sustained:
            line["sustained"] = sustained
        if legs:
            line["configs"] = legs
        if shard_io:
            line["shard_io"] = shard_io
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sdpa_baseline(CONFIGS["C2"][1], CONFIGS["C2"][3], CONFIGS["C2"][4])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
