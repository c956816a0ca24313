#!/usr/bin/env python
"""BASELINE.json configs[3]: batched NTT, N = 2^16, 2^28 coefficients in total (4096
polynomials), sharded across the GPUs of one box with NO communication (strong scaling:
total work fixed, each rank transforms 4096/G polynomials).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port P tools/bench_cfg4_sharded.py [--reps 20]

One JSON line from rank 0: whole-job polys/s (device-timed, max over ranks), algorithmic
GB/s (8 N bytes per polynomial) per GPU against the measured HBM peak, and sampled
parity against the oracle on every rank.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntt_aie_b200 as nt  # noqa: E402
from ntt_aie_b200.fourstep import shard_batch  # noqa: E402

Q, LOGN, TOTAL = 469762049, 16, 4096


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << LOGN
    b0, b1 = shard_batch(TOTAL, world, rank)
    batch = b1 - b0
    roots = nt.make_roots(n, Q, 3)
    gen = torch.Generator(device="cuda").manual_seed(0x5EED0016 + rank)
    d_in = torch.randint(0, Q, (batch, n), dtype=torch.int32, device="cuda", generator=gen)
    d_out = torch.empty_like(d_in)
    plan = nt.Plan(LOGN, Q, roots, device=local)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        plan.gs(d_in, d_out, batch)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        plan.gs(d_in, d_out, batch)
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) / args.reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    import oracle
    idx = [0, batch - 1, batch // 2]
    ok = bool(np.array_equal(d_out[idx].cpu().numpy(),
                             oracle.ntt_gs(d_in[idx].cpu().numpy(), roots, Q)))
    okt = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak = 6650.0
        gbs_per_gpu = batch * n * 8 / (ms * 1e-3) / 1e9
        print(json.dumps({
            "config": f"cfg4 batched NTT N=2^16, 4096 polys (2^28 coefficients) over {world} GPU(s), no communication",
            "n_gpus": world, "polys_per_gpu": batch, "ms": ms, "polys_per_s": TOTAL / (ms * 1e-3),
            "butterflies_per_s": TOTAL * (n // 2) * LOGN / (ms * 1e-3),
            "algorithmic_GBps_per_gpu": gbs_per_gpu, "frac_of_measured_hbm_per_gpu": gbs_per_gpu / peak,
            "scaling": "strong", "kernel_path": plan.last_path, "bit_exact_sampled_all_ranks": bool(okt.item())}),
            flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
