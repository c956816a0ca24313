#!/usr/bin/env python
"""Profile tooling parity: regenerates the CSV tables that correspond to the
reference's profile/kerneltime/*.csv and profile/plot_efficiency.py, from this repo's
measured JSON lines (CUDA-event timings written by tools/bench_configs.py and bench.py).

    python profiles/make_tables.py        # writes profiles/kerneltime_b200.csv, efficiency_b200.csv

kerneltime_b200.csv   N, microseconds per transform (batched: launch time / batch) next to the
                      reference's per-launch kernel times on the AIE (profile/kerneltime/aie.csv)
                      and on their A100 (profile/kerneltime/gpu.csv) as quoted in BASELINE.md.
efficiency_b200.csv   N, achieved fraction of the measured HBM roofline (this repo's metric) and
                      the reference's "efficiency" (5.5*N*log2N ops over 88 GOPS AIE / 4280 GOPS
                      A100, profile/plot_efficiency.py:25-27,44-46) for context.
matplotlib is not in this image, so the plots themselves are left to the reader.
"""
import json
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# reference numbers (microseconds per single-transform launch)
AIE_US = {512: 8.863, 1024: 10.676, 2048: 14.375, 4096: 22.065}
A100_US = {256: 12.004, 512: 13.497, 1024: 16.365, 2048: 21.510, 4096: 19.276, 8192: 21.179,
           16384: 24.203, 32768: 31.337, 65536: 45.942, 131072: 81.350}


def main():
    rows = {}
    path = os.path.join(HERE, "r1_configs_0_2_3.jsonl")
    for line in open(path):
        d = json.loads(line)
        if d["config"].startswith("ntt forward GS"):
            n = 1 << d["logn"]
            rows[n] = {"us_per_transform": 1e6 / d["polys_per_s"], "frac": d["frac_of_measured_hbm"],
                       "bfly_per_s": d["butterflies_per_s"], "path": d["kernel_path"]}
    with open(os.path.join(HERE, "kerneltime_b200.csv"), "w") as f:
        f.write("N,b200_us_per_transform_batched,b200_kernel_path,ref_aie_us_per_launch,ref_a100_us_per_launch\n")
        for n in sorted(rows):
            f.write(f"{n},{rows[n]['us_per_transform']:.5f},{rows[n]['path']},"
                    f"{AIE_US.get(n, '')},{A100_US.get(n, '')}\n")
    with open(os.path.join(HERE, "efficiency_b200.csv"), "w") as f:
        f.write("N,b200_frac_of_measured_hbm_roofline,b200_butterflies_per_s,"
                "ref_aie_efficiency_of_88GOPS,ref_a100_efficiency_of_4280GOPS\n")
        for n in sorted(rows):
            ops = 5.5 * math.log2(n) * n
            aie = ops / (1000 * AIE_US[n]) / 88 if n in AIE_US else ""
            a100 = ops / (1000 * A100_US[n]) / 4280 if n in A100_US else ""
            f.write(f"{n},{rows[n]['frac']:.4f},{rows[n]['bfly_per_s']:.4g},"
                    f"{aie if aie == '' else round(aie, 4)},{a100 if a100 == '' else round(a100, 5)}\n")
    print(open(os.path.join(HERE, "kerneltime_b200.csv")).read())
    print(open(os.path.join(HERE, "efficiency_b200.csv")).read())


if __name__ == "__main__":
    main()
