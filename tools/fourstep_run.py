#!/usr/bin/env python
"""Single large transform over the GPUs of one box (BASELINE.json configs[4]).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port P tools/fourstep_run.py --logn 26 [--verify] [--steps K] [--fused] [--generated]

One process per GPU, NCCL for the one all-to-all transpose (ntt-aie_b200/fourstep.py), or
-- with --fused -- the transposes fused into the passes' NVLink peer stores.  Prints one
JSON line from rank 0: device-timed (CUDA events, max over ranks) milliseconds per
transform for the transposed-order and the natural-order variants, the all-to-all share,
and parity of the WHOLE vector:
  --verify          against the CPU golden (oracle; test infrastructure) on rank 0, both orders;
  --verify-digest   against the committed digest of the reference's own golden output
                    (tests/golden/large_digests.npz; logn 22/24/26/27, reference table only)
                    -- no CPU transform, used by bench.py.
`run()` is the same thing as a function (bench.py appends its result to the bench line).
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
G = 3


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run(logn=26, steps=5, warmup=2, verify=False, verify_digest=False, fused=False,
        arbitrary_table=False, generated=False, detail=True):
    """Runs on every rank of an initialised process group (or alone); rank 0 gets the dict."""
    import ntt_aie_b200 as nt
    from ntt_aie_b200.fourstep import FourStepNTT
    from tools.digest import as_unsigned, digest_numpy, digest_torch

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    local = torch.cuda.current_device()
    n = 1 << logn
    s = n // world
    t0 = time.perf_counter()
    table = None
    if arbitrary_table:
        table = np.random.default_rng(99).integers(0, Q, n, dtype=np.int32)
    elif not generated or verify:
        table = nt.make_roots(n, Q, G)                  # reference convention (src/test.cpp:27-32)
    seed = 0x5EED0026 + logn                            # the input of tests/golden/make_golden_large.py
    a = np.random.default_rng(seed).integers(0, Q, n, dtype=np.int32)   # same vector on every rank
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    if generated:
        w = nt.powmod(G, (Q - 1) // n, Q)
        plan = FourStepNTT(logn, Q, None, rank, world, device=local, fused=fused,
                           generated=(nt.GEN_POWERS, w))
    else:
        plan = FourStepNTT(logn, Q, table, rank, world, device=local, fused=fused)
    torch.cuda.synchronize()
    t_plan = time.perf_counter() - t0
    shard0 = torch.from_numpy(a[rank * s:(rank + 1) * s].copy()).cuda()
    shard = torch.empty_like(shard0)
    scratch = torch.empty_like(shard0)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(natural):
        for _ in range(warmup):
            shard.copy_(shard0)
            plan.forward(shard, scratch, natural_order=natural)
        tot = 0.0
        for _ in range(steps):
            shard.copy_(shard0)
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.forward(shard, scratch, natural_order=natural)
            e1.record()
            sync()
            tot += e0.elapsed_time(e1)
        return reduce_max(tot / steps)

    ms_dev = timed(False)
    ms_nat = timed(True)

    def timed_fn(fn):
        for _ in range(2):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        return reduce_max(e0.elapsed_time(e1) / steps)

    a2a_ms = ms_local = ms_cross = None
    if detail:
        # the all-to-all alone (same buffers), for the NVLink roofline
        if world > 1:
            a2a_ms = timed_fn(lambda: dist.all_to_all_single(scratch, shard))
        # the local phases alone (HBM roofline of the passes)
        eng = plan.engine
        ms_local = timed_fn(lambda: eng.local_full(shard))
        if world > 1:
            logc = plan.logs - (world.bit_length() - 1)
            ms_cross = timed_fn(lambda: eng.cross_stages(scratch, logc, plan.logs))

    ok = ok_dev = None
    cpu_s = None
    if verify or verify_digest:
        want = None
        if verify and rank == 0:
            import oracle
            t1 = time.perf_counter()
            want = oracle.ntt_gs(a, table, Q)
            cpu_s = time.perf_counter() - t1
        c = s // world
        # natural order: rank r holds NTT(a)[r*S, (r+1)*S)
        shard.copy_(shard0)
        res = plan.forward(shard, scratch, natural_order=True).clone()
        d_nat = digest_torch(res, offset=rank * s)
        # transposed order: rank k holds out_k[r][c] = NTT(a)[r*S + k*S/G + c]
        shard.copy_(shard0)
        res_t = plan.forward(shard, scratch, natural_order=False).clone()
        d_tr = torch.zeros(2, dtype=torch.int64, device="cuda")
        for r in range(world):
            d_tr += digest_torch(res_t[r * c:(r + 1) * c], offset=r * s + rank * c)
        if world > 1:
            dist.all_reduce(d_nat)
            dist.all_reduce(d_tr)
        if verify_digest and not arbitrary_table:
            g = np.load(os.path.join(ROOT, "tests", "golden", "large_digests.npz"))
            if f"digest_{logn}" in g and rank == 0:
                ref = tuple(int(v) for v in g[f"digest_{logn}"])
                ok, ok_dev = as_unsigned(d_nat) == ref, as_unsigned(d_tr) == ref
        if verify and rank == 0:
            ref = digest_numpy(want)
            ok, ok_dev = as_unsigned(d_nat) == ref, as_unsigned(d_tr) == ref
            # and word for word on this rank's own part
            ok = ok and bool(np.array_equal(res.cpu().numpy(), want[:s]))
    line = None
    if rank == 0:
        bfly = (n // 2) * logn
        sent = (world - 1) / world * 4 * s if world > 1 else 0   # bytes each GPU sends per transpose
        line = {
            "workload": f"single four-step NTT N=2^{logn}, q={Q}, {world} GPU(s)",
            "n_gpus": world, "logn": logn,
            "exchange": "fused peer stores (symmetric memory)" if fused else "NCCL all_to_all_single",
            "tables": "generated on device" if generated else "host table",
            "ms_transposed_order": ms_dev, "ms_natural_order": ms_nat,
            "butterflies_per_s_transposed": bfly / (ms_dev * 1e-3),
            "butterflies_per_s_natural": bfly / (ms_nat * 1e-3),
            "exchange_bytes_per_gpu_per_dir": sent,
            "all_to_all_ms": a2a_ms,
            "all_to_all_GBps_per_gpu_per_dir": (sent / (a2a_ms * 1e-3) / 1e9) if a2a_ms else None,
            "nvlink_peak_GBps_per_dir": 900.0, "nvlink_measured_peer_copy_GBps": 770.0,
            "local_stages_ms": ms_local, "cross_stages_ms": ms_cross,
            "local_stages_frac_of_measured_hbm": (8 * s / (ms_local * 1e-3) / 1e9 / hbm_peak()) if ms_local else None,
            "cross_stages_frac_of_measured_hbm": (8 * s / (ms_cross * 1e-3) / 1e9 / hbm_peak()) if ms_cross else None,
            "host_table_and_input_s": t_host, "plan_build_s": t_plan,
            "bit_exact_vs_golden": ok, "bit_exact_transposed_order": ok_dev,
            "verified_against": ("cpu golden (oracle)" if verify else
                                 "digest of the reference's golden output (tests/golden/large_digests.npz)"
                                 if verify_digest else None),
        }
        if cpu_s is not None:
            line["cpu_golden_s_single_thread"] = cpu_s
    plan.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=26)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--verify-digest", action="store_true")
    ap.add_argument("--fused", action="store_true",
                    help="transposes fused into the passes' NVLink stores (symmetric memory)")
    ap.add_argument("--arbitrary-table", action="store_true",
                    help="random table instead of the reference's w^i (table-driven check)")
    ap.add_argument("--generated", action="store_true",
                    help="per-rank tables generated on the device (nothing shipped)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = run(args.logn, args.steps, args.warmup, args.verify, args.verify_digest, args.fused,
               args.arbitrary_table, args.generated)
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
