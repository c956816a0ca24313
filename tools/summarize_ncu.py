#!/usr/bin/env python
"""Summarise ncu captures (gpurun_out/prof_*.ncu-rep + a launch-list CSV) into the text
files kept under profiles/.  Run where ncu is installed (no GPU needed):
    python tools/summarize_ncu.py gpurun_out profiles/r1_secondary_kernels_ncu.txt \
        [--launches gpurun_out/launches_configs.csv] kernel_a kernel_b ...
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
        'smsp__sass_inst_executed_op_tmem_ldt.sum', 'smsp__sass_inst_executed_op_tmem_stt.sum',
        'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct']


def main():
    args = sys.argv[1:]
    src, out_path = args[0], args[1]
    launches = None
    if "--launches" in args:
        i = args.index("--launches")
        launches = args[i + 1]
        del args[i:i + 2]
    kernels = args[2:]
    lines = []
    for k in kernels:
        raw = subprocess.run(['ncu', '-i', f'{src}/prof_{k}.ncu-rep', '--page', 'raw', '--csv'],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        lines.append(f"== {d['Kernel Name'][0][:120]}")
        for kk in KEYS:
            if kk in d:
                lines.append(f"   {kk:70s} {d[kk][0]} {d[kk][1]}")
        st = [(float(d[h][0]), h) for h in d
              if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
        lines.append("   top stalls: " + ", ".join(
            f"{h.split('issue_stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in sorted(st, reverse=True)[:6]))
    if launches:
        rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
        t, n = collections.Counter(), collections.Counter()
        for r in rows:
            name = r[4].split('(')[0].replace('void ', '').replace('nttb200::', '')[:52]
            t[name] += float(r[-1])
            n[name] += 1
        tot = sum(t.values())
        lines.append(f"\n== launch list {launches} (ncu: cold-cache, serialised -- compare shares)")
        for k, v in t.most_common(14):
            lines.append(f"   {v / 1e3:10.1f} us {100 * v / tot:5.1f}%  n={n[k]:4d}  mean {v / n[k] / 1e3:8.1f} us  {k}")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
