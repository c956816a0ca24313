#!/usr/bin/env python
"""Single large transform over the GPUs of one box (BASELINE.json configs[4]).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port P tools/fourstep_run.py --logn 26 [--verify] [--steps K]

One process per GPU, NCCL for the one all-to-all transpose (ntt-aie_b200/fourstep.py).
Prints one JSON line from rank 0: device-timed (CUDA events, max over ranks) seconds per
transform for the transposed-order and the natural-order variants, the all-to-all share,
and -- with --verify -- bit-exact parity of the whole vector against the CPU golden
(oracle; test infrastructure) on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

Q = 469762049


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=26)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--fused", action="store_true",
                    help="transposes fused into the passes' NVLink stores (symmetric memory)")
    ap.add_argument("--arbitrary-table", action="store_true",
                    help="random table instead of the reference's w^i (table-driven check)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import ntt_aie_b200 as nt
    from ntt_aie_b200.fourstep import FourStepNTT

    n = 1 << args.logn
    s = n // world
    t0 = time.perf_counter()
    if args.arbitrary_table:
        table = np.random.default_rng(99).integers(0, Q, n, dtype=np.int32)
    else:
        table = nt.make_roots(n, Q, 3)                  # reference convention (src/test.cpp:27-32)
    t_table = time.perf_counter() - t0
    rng = np.random.default_rng(0x5EED0026)
    a = rng.integers(0, Q, n, dtype=np.int32)            # same vector on every rank (seeded)
    plan = FourStepNTT(args.logn, Q, table, rank, world, device=local, fused=args.fused)
    shard0 = torch.from_numpy(a[rank * s:(rank + 1) * s].copy()).cuda()
    shard = torch.empty_like(shard0)
    scratch = torch.empty_like(shard0)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(natural):
        for _ in range(args.warmup):
            shard.copy_(shard0)
            plan.forward(shard, scratch, natural_order=natural)
        tot = 0.0
        for _ in range(args.steps):
            shard.copy_(shard0)
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.forward(shard, scratch, natural_order=natural)
            e1.record()
            sync()
            tot += e0.elapsed_time(e1)
        t = torch.tensor([tot / args.steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_dev = timed(False)
    ms_nat = timed(True)

    # the all-to-all alone (same buffers), for the NVLink roofline
    a2a_ms = None
    if world > 1 and not args.fused:
        for _ in range(2):
            dist.all_to_all_single(scratch, shard)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            dist.all_to_all_single(scratch, shard)
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        a2a_ms = float(t.item())

    # the local phases alone (HBM roofline of the passes): step 1 on the shard, step 3 on scratch
    def timed_fn(fn):
        for _ in range(2):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = plan.engine
    ms_local = timed_fn(lambda: eng.local_full(shard))
    ms_cross = None
    if world > 1:
        logc = plan.logs - (world.bit_length() - 1)
        ms_cross = timed_fn(lambda: eng.cross_stages(scratch, logc, plan.logs))

    ok = None
    if args.verify:
        shard.copy_(shard0)
        res = plan.forward(shard, scratch, natural_order=True).clone()
        parts = [torch.empty_like(res) for _ in range(world)] if rank == 0 else None
        if world > 1:
            dist.gather(res, parts, dst=0)
        else:
            parts = [res]
        if rank == 0:
            import oracle
            got = torch.cat(parts).cpu().numpy()
            t1 = time.perf_counter()
            want = oracle.ntt_gs(a, table, Q)
            ok = bool(np.array_equal(got, want))
            cpu_s = time.perf_counter() - t1
    if rank == 0:
        bfly = (n // 2) * args.logn
        sent = (world - 1) / world * 4 * s if world > 1 else 0   # bytes each GPU sends per a2a
        line = {
            "workload": f"single four-step NTT N=2^{args.logn}, q={Q}, {world} GPU(s)",
            "n_gpus": world, "logn": args.logn,
            "exchange": "fused peer stores (symmetric memory)" if args.fused else "NCCL all_to_all_single",
            "ms_transposed_order": ms_dev, "ms_natural_order": ms_nat,
            "butterflies_per_s_transposed": bfly / (ms_dev * 1e-3),
            "butterflies_per_s_natural": bfly / (ms_nat * 1e-3),
            "all_to_all_ms": a2a_ms,
            "all_to_all_GBps_per_gpu_per_dir": (sent / (a2a_ms * 1e-3) / 1e9) if a2a_ms else None,
            "nvlink_peak_GBps_per_dir": 900.0, "nvlink_measured_peer_copy_GBps": 770.0,
            "local_stages_ms": ms_local, "cross_stages_ms": ms_cross,
            "local_stages_algorithmic_GBps_per_gpu": 8 * s / (ms_local * 1e-3) / 1e9,
            "local_stages_frac_of_measured_hbm": 8 * s / (ms_local * 1e-3) / 1e9 / hbm_peak(),
            "cross_stages_frac_of_measured_hbm": (8 * s / (ms_cross * 1e-3) / 1e9 / hbm_peak()) if ms_cross else None,
            "table_build_s": t_table, "bit_exact_vs_golden": ok,
        }
        if args.verify:
            line["cpu_golden_s_single_thread"] = cpu_s
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
