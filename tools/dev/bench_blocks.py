"""Run bench.py's polymul / cfg4 blocks alone and print the per-launch times (debugging aid)."""
import json, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
import torch
import ntt_aie_b200 as nt
ctx = bench.Ctx()
peak = 6539.5
orig = bench.time_launches
def spy(ctx_, fn, reps, warmup=3):
    total, ms = orig(ctx_, fn, reps, warmup)
    print("  launches ms:", [round(x, 4) for x in ms], flush=True)
    return total, ms
bench.time_launches = spy
which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
if which == "cfg4":
    print(json.dumps(bench.block_cfg4_strong(ctx, nt, peak, 10))[:400])
else:
    for r in bench.block_polymul_sweep(ctx, nt, peak, 10):
        print(r["logn"], round(r["frac_of_measured_hbm"], 4), r["kernel_path"])
