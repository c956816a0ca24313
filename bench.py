#!/usr/bin/env python
"""bench.py -- batched NTT throughput on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  Workload
(BASELINE.json configs[1]): 65,536 independent polynomials, N=4096, 32-bit prime
q = 469762049 (7*2^26+1), twiddle table in the reference's convention
(roots[i] = w^i, w = 3^((q-1)/N), src/test.cpp:27-32), golden GS network
(src/test.cpp:34-60).  Per-GPU work is fixed as N grows (weak scaling, batched
polynomials shard with no communication).

Numbers on the one JSON line rank 0 prints:
  value      polys/s, whole job, inputs resident in HBM, device-timed (CUDA events on
             the launching stream, max over ranks);
  e2e        the same metric through the host-buffer entry point nttb200_gs_host
             (pinned host buffers, H2D + kernel + D2H every step inside the timed region);
  roofline   algorithmic bytes (8 B per coefficient: 4 read + 4 written) per launch /
             mean launch duration, against MEASURED_PEAKS.json hbm_gbs;
  cpu_baseline  the reference's own golden ntt() (oracle/_ref, compiled from
             /root/reference/src/test.cpp:15-60) on the host cores, bounded sample.

--impl reference times only that CPU golden (all host threads) and prints the same
line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOGN = 12
N = 1 << LOGN
Q = 469762049
G = 3
BATCH = 65536
METRIC = "batched NTT polys/sec (N=4096, 32-bit q)"
WORKLOAD = "batched forward NTT: 65,536 independent polynomials, N=4096, q=469762049, 1xB200"
BYTES_PER_POLY = 8 * N                 # algorithmic: 4 B read + 4 B written per coefficient
BFLY_PER_POLY = (N // 2) * LOGN
FALLBACK_HBM_GBS = 6650.0              # B200_PROFILING.md fallback


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per launch of the dominant kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("fused_gs4096_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------- CPU golden
def cpu_golden_rate(sample_polys: int, threads: int, repeats: int = 2):
    """polys/s of the reference's golden ntt() on `threads` host threads."""
    import oracle
    roots = oracle.make_roots(N, Q, G)
    rng = np.random.default_rng(0x5EED0001)
    a = rng.integers(0, Q, (sample_polys, N), dtype=np.int32)
    if oracle.have_ref():
        kind, fn = "reference", oracle.ref_ntt_batch_inplace
    else:
        oracle.build()
        kind, fn = "port", oracle.ntt_gs_batch_inplace
    best = None
    for _ in range(repeats):
        work = a.copy()
        t0 = time.perf_counter()
        fn(work, roots, Q, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample_polys / best, kind, best


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    sample = max(threads * 256, 4096)         # polynomials per step (~1 s of CPU work each)
    import oracle
    roots = oracle.make_roots(N, Q, G)
    if oracle.have_ref():
        kind, fn = "reference", oracle.ref_ntt_batch_inplace
    else:
        oracle.build()
        kind, fn = "port", oracle.ntt_gs_batch_inplace
    rng = np.random.default_rng(0x5EED0001)
    a = rng.integers(0, Q, (sample, N), dtype=np.int32)
    work = a.copy()
    for _ in range(args.warmup):
        np.copyto(work, a)
        fn(work, roots, Q, threads)
    total = 0.0
    for _ in range(args.steps):
        np.copyto(work, a)                    # untimed: restore canonical inputs
        t0 = time.perf_counter()
        fn(work, roots, Q, threads)
        total += time.perf_counter() - t0
    value = sample * args.steps / total
    sample_txt = (f"{sample} polys/step x {args.steps} steps of the same workload "
                  f"(N={N}, q={Q}) on {threads} host threads, g++ -O2")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "polys/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "logn": LOGN, "q": Q, "polys_per_step": sample,
                   "note": "reference's own CPU golden ntt() (src/test.cpp:34-60); host only"},
        "butterflies_per_s": value * BFLY_PER_POLY,
        "cpu_baseline": {"value": value, "unit": "polys/s", "cores": threads, "kind": kind,
                         "sample": sample_txt},
        "e2e": {"value": value, "unit": "polys/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------ GPU arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import ntt_aie_b200 as nt
    nt.load_library()                          # fails loudly if the .so is missing

    batch = args.batch
    roots = nt.make_roots(N, Q, G)
    plan = nt.Plan(LOGN, Q, roots, device=local_rank)
    gen = torch.Generator(device="cuda").manual_seed(0x5EED0001 + rank)
    d_in = torch.randint(0, Q, (batch, N), dtype=torch.int32, device="cuda", generator=gen)
    d_out = torch.empty_like(d_in)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------
    for _ in range(args.warmup):
        plan.gs(d_in, d_out, batch, -1, stream)
    barrier()
    path = plan.last_path
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = nt.kernel_launches()
    barrier()
    ev[0].record(stream)
    for k in range(args.steps):
        plan.gs(d_in, d_out, batch, -1, stream)
        ev[k + 1].record(stream)
    barrier()
    launches = nt.kernel_launches() - launches0
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * batch * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the host-buffer entry point -----------------------------
    e2e_steps = max(3, min(args.steps, 10))
    if args.no_e2e:
        e2e_steps = 0
    h_in = torch.empty((batch, N), dtype=torch.int32, pin_memory=True)
    h_out = torch.empty((batch, N), dtype=torch.int32, pin_memory=True)
    e2e_value, same = None, None
    if e2e_steps:
        h_in.copy_(d_in)
        plan.gs_host(h_in, h_out, batch)      # warm-up (allocates the staging ring)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.gs_host(h_in, h_out, batch)  # synchronous: returns when h_out is complete
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = world * batch * e2e_steps / float(te.item())
        # cheap integrity check of the e2e result against the device-resident one
        same = bool(torch.equal(h_out[:64].cuda(), d_out[:64])) and bool(
            torch.equal(h_out[-64:].cuda(), d_out[-64:]))

    if rank == 0:
        peak, peak_src = measured_peaks()
        mean_launch_ms = statistics.mean(per_launch_ms)
        achieved = batch * BYTES_PER_POLY / (mean_launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "polys/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "logn": LOGN, "q": Q, "polys_per_gpu": batch,
                       "table": "reference convention roots[i]=w^i (src/test.cpp:27-32)",
                       "kernel_path": path, "sharding": f"batch x{world}, no communication",
                       "l2": "inputs larger than L2 (1 GiB in + 1 GiB out per step)"},
            "butterflies_per_s": value * BFLY_PER_POLY,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "polys/s",
                    "h2d_bytes_per_step": batch * N * 4, "d2h_bytes_per_step": batch * N * 4,
                    "steps": e2e_steps, "api": "nttb200_gs_host (pinned host buffers)",
                    "matches_device_result": same},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_per_launch(),
                         "peak_source": peak_src, "kernel": path,
                         "algorithmic_bytes_per_launch": batch * BYTES_PER_POLY,
                         "mean_launch_ms": mean_launch_ms,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            sample = BATCH                        # the full workload: ~10 core-seconds
            rate, kind, secs = cpu_golden_rate(sample, threads)
            rate1, _, secs1 = cpu_golden_rate(1024, 1, repeats=1)
            line["cpu_baseline"] = {
                "value": rate, "unit": "polys/s", "cores": threads, "kind": kind,
                "sample": f"{sample} polys of the same workload (N={N}, q={Q}), best of 2, "
                          f"{secs:.3f} s wall on {threads} threads",
                "single_thread_polys_per_s": rate1}
        print(json.dumps(line), flush=True)
    plan.close()
    if distributed:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=BATCH, help="polynomials per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
