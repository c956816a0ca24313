#!/usr/bin/env python
"""Profile tooling parity: regenerates the tables AND the plots that correspond
to the reference's profile/ directory from this repo's measured JSON lines (CUDA-event and
wall-clock timings written by tools/bench_configs.py).

    python profiles/make_tables.py [profiles/r2_configs.jsonl]

  kerneltime_b200.csv / kerneltime.svg   kernel time per transform against N, as
        profile/plot_kerneltime.py draws it: one single transform on the B200 (device time of
        one launch), the batched per-transform time, and the reference's per-launch kernel
        times on the AIE (profile/kerneltime/aie.csv) and on their A100 (profile/kerneltime/gpu.csv)
  exectime_b200.csv / exectime.svg       host wall time per single-transform call (trimmed mean
        dropping min and max, profile/plot_exectime.py:27-29) against N, next to the reference's
        16-tile NPU wall time where BASELINE.md quotes it
  efficiency_b200.csv / efficiency.svg   achieved fraction of the measured HBM roofline (this
        repo's metric), GS and CT, and the reference's "efficiency" (5.5*N*log2N ops over 88 GOPS
        AIE / 4280 GOPS A100, profile/plot_efficiency.py:25-27,44-46) for context
matplotlib is not in this image, so the plots are written as plain SVG by the few lines below.
"""
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
# reference numbers (microseconds per single-transform launch)
AIE_US = {512: 8.863, 1024: 10.676, 2048: 14.375, 4096: 22.065}
A100_US = {256: 12.004, 512: 13.497, 1024: 16.365, 2048: 21.510, 4096: 19.276, 8192: 21.179,
           16384: 24.203, 32768: 31.337, 65536: 45.942, 131072: 81.350}
AIE_WALL_US = {2048: 279.0}   # host wall time per launch, the one point BASELINE.md quotes
COLORS = ["#1f77b4", "#2ca02c", "#d62728", "#9467bd", "#ff7f0e", "#8c564b"]


def svg_plot(path, title, xlabel, ylabel, series, logy=False):
    """series: list of (label, {N: value}); x axis log2(N)."""
    w, h, ml, mr, mt, mb = 860, 520, 90, 30, 50, 70
    xs = sorted({n for _, d in series for n in d})
    ys = [v for _, d in series for v in d.values() if v is not None and (v > 0 or not logy)]
    x0, x1 = math.log2(xs[0]), math.log2(xs[-1])
    fy = (lambda v: math.log10(v)) if logy else (lambda v: v)
    y0 = fy(min(ys)) if logy else 0.0
    y1 = fy(max(ys))
    y1 += 0.08 * (y1 - y0 if y1 > y0 else 1)

    def px(n):
        return ml + (math.log2(n) - x0) / max(x1 - x0, 1e-9) * (w - ml - mr)

    def py(v):
        return h - mb - (fy(v) - y0) / max(y1 - y0, 1e-9) * (h - mt - mb)

    o = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{w}" height="{h}" font-family="sans-serif">',
         f'<rect width="{w}" height="{h}" fill="white"/>',
         f'<text x="{w / 2}" y="28" text-anchor="middle" font-size="18">{title}</text>',
         f'<text x="{w / 2}" y="{h - 18}" text-anchor="middle" font-size="15">{xlabel}</text>',
         f'<text x="22" y="{h / 2}" text-anchor="middle" font-size="15" '
         f'transform="rotate(-90 22 {h / 2})">{ylabel}</text>',
         f'<line x1="{ml}" y1="{h - mb}" x2="{w - mr}" y2="{h - mb}" stroke="black"/>',
         f'<line x1="{ml}" y1="{mt}" x2="{ml}" y2="{h - mb}" stroke="black"/>']
    for n in xs:
        o.append(f'<line x1="{px(n):.1f}" y1="{h - mb}" x2="{px(n):.1f}" y2="{h - mb + 5}" stroke="black"/>')
        o.append(f'<text x="{px(n):.1f}" y="{h - mb + 20}" text-anchor="middle" font-size="12">'
                 f'2^{int(math.log2(n))}</text>')
    if logy:
        ticks = [10 ** e for e in range(math.floor(y0), math.ceil(y1) + 1)]
    else:
        step = 10 ** math.floor(math.log10(max(y1, 1e-9)))
        step = step / 2 if y1 / step < 4 else step
        ticks = [k * step for k in range(int(y1 / step) + 1)]
    for t in ticks:
        if t <= 0 and logy:
            continue
        if not (y0 - 1e-9 <= fy(t) <= y1 + 1e-9):
            continue
        o.append(f'<line x1="{ml}" y1="{py(t):.1f}" x2="{w - mr}" y2="{py(t):.1f}" stroke="#dddddd"/>')
        o.append(f'<text x="{ml - 8}" y="{py(t) + 4:.1f}" text-anchor="end" font-size="12">{t:g}</text>')
    for k, (label, d) in enumerate(series):
        c = COLORS[k % len(COLORS)]
        pts = [(px(n), py(v)) for n, v in sorted(d.items()) if v is not None and (v > 0 or not logy)]
        if len(pts) > 1:
            o.append('<polyline fill="none" stroke="%s" stroke-width="2" points="%s"/>'
                     % (c, " ".join(f"{x:.1f},{y:.1f}" for x, y in pts)))
        for x, y in pts:
            o.append(f'<circle cx="{x:.1f}" cy="{y:.1f}" r="4" fill="{c}"/>')
        o.append(f'<rect x="{ml + 14}" y="{mt + 8 + 20 * k}" width="14" height="4" fill="{c}"/>')
        o.append(f'<text x="{ml + 34}" y="{mt + 14 + 20 * k}" font-size="13">{label}</text>')
    o.append("</svg>")
    open(path, "w").write("\n".join(o) + "\n")


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "r2_configs.jsonl")
    gs, ct, single, wall, mul = {}, {}, {}, {}, {}
    for line in open(src):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        cfg = d.get("config", "")
        if cfg.startswith("ntt forward GS"):
            gs[1 << d["logn"]] = d
        elif cfg.startswith("ntt forward CT"):
            ct[1 << d["logn"]] = d
        elif cfg.startswith("exectime"):
            single[1 << d["logn"]] = d["device_kernel_us_trimmed_mean"]
            wall[1 << d["logn"]] = d["host_call_us_trimmed_mean"]
        elif "negacyclic polymul" in cfg and "rns" not in cfg:
            mul[1 << d["logn"]] = d
    batched = {n: 1e6 / d["polys_per_s"] for n, d in gs.items()}
    with open(os.path.join(HERE, "kerneltime_b200.csv"), "w") as f:
        f.write("N,b200_us_single_transform_kernel,b200_us_per_transform_batched,b200_kernel_path,"
                "ref_aie_us_per_launch,ref_a100_us_per_launch\n")
        for n in sorted(set(batched) | set(single)):
            f.write(f"{n},{single.get(n, '')},{batched.get(n, '')},{gs[n]['kernel_path'] if n in gs else ''},"
                    f"{AIE_US.get(n, '')},{A100_US.get(n, '')}\n")
    with open(os.path.join(HERE, "exectime_b200.csv"), "w") as f:
        f.write("N,b200_host_call_us_trimmed_mean,ref_aie_16tile_wall_us\n")
        for n in sorted(wall):
            f.write(f"{n},{wall[n]:.2f},{AIE_WALL_US.get(n, '')}\n")
    eff_aie = {n: 5.5 * math.log2(n) * n / (1000 * us) / 88 for n, us in AIE_US.items()}
    eff_a100 = {n: 5.5 * math.log2(n) * n / (1000 * us) / 4280 for n, us in A100_US.items()}
    with open(os.path.join(HERE, "efficiency_b200.csv"), "w") as f:
        f.write("N,b200_gs_frac_of_measured_hbm_roofline,b200_ct_frac,b200_polymul_frac,"
                "b200_butterflies_per_s,ref_aie_efficiency_of_88GOPS,ref_a100_efficiency_of_4280GOPS\n")
        for n in sorted(set(gs) | set(ct) | set(mul)):
            f.write(f"{n},{gs[n]['frac_of_measured_hbm'] if n in gs else '':.4},"
                    f"{ct[n]['frac_of_measured_hbm'] if n in ct else '':.4},"
                    f"{mul[n]['frac_of_measured_hbm'] if n in mul else '':.4},"
                    f"{gs[n]['butterflies_per_s'] if n in gs else '':.4},"
                    f"{round(eff_aie[n], 4) if n in eff_aie else ''},"
                    f"{round(eff_a100[n], 5) if n in eff_a100 else ''}\n")
    svg_plot(os.path.join(HERE, "kerneltime.svg"), "Kernel time per transform", "Data size",
             "Kernel Time (us)",
             [("B200, one transform per launch", single), ("B200, batched (launch / batch)", batched),
              ("Ryzen AI Engine (reference)", AIE_US), ("A100 (reference)", A100_US)], logy=True)
    svg_plot(os.path.join(HERE, "exectime.svg"), "Host wall time per single-transform call", "Data size",
             "Execution Time (us)",
             [("B200, nttb200_gs_host", wall), ("Ryzen AI Engine, 16 tiles (reference)", AIE_WALL_US)])
    svg_plot(os.path.join(HERE, "efficiency.svg"), "Efficiency", "Data size", "fraction of roofline",
             [("B200 GS: fraction of measured HBM roofline", {n: d["frac_of_measured_hbm"] for n, d in gs.items()}),
              ("B200 CT", {n: d["frac_of_measured_hbm"] for n, d in ct.items()}),
              ("B200 negacyclic product (12N bytes)", {n: d["frac_of_measured_hbm"] for n, d in mul.items()}),
              ("Ryzen AI Engine: ops / 88 GOPS (reference)", eff_aie),
              ("A100: ops / 4280 GOPS (reference)", eff_a100)])
    for name in ("kerneltime_b200.csv", "exectime_b200.csv", "efficiency_b200.csv"):
        print(open(os.path.join(HERE, name)).read())


if __name__ == "__main__":
    main()
