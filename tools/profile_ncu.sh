#!/bin/bash
# ncu recipe behind profiles/ (run on a B200 box: `gpurun -- bash tools/profile_ncu.sh`).
# Successor of the reference's trace tooling (event0/event1 markers + scripts/parse_trace.py +
# profile/plot_*.py): (1) the launch list of the bench command (per-launch device time; compare
# SHARES, the numbers are cold-cache and serialised), (2) one `--set full` capture per kernel
# family.  Every command is profiled only after the same command exited 0 without ncu.
# Outputs land in gpurun_out/; summarise with tools/summarize_ncu.py (no GPU needed).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/launches_bench.csv $CMD > /dev/null 2>&1
cap() {  # cap <target> <kernel regex> <report name>
  python tools/dev/ncu_targets.py $1 > gpurun_out/plain_$1.log 2>&1 || { echo "plain $1 failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f \
      -o gpurun_out/prof_$3 python tools/dev/ncu_targets.py $1 > gpurun_out/ncu_$3.log 2>&1
}
cap headline fused_gs4096 fused_gs4096
cap polymul polymul4096 polymul4096
cap poly15 polyt_gs polyt_gs15
cap tilecol16 tilecol_gs tilecol_gs16
cap ct4096 tile_ct_h tile_ct_h
cap ct15 polyt_ct polyt_ct15
cap ct16 tilecol_ct tilecol_ct16   # (all seven reports exceed the 64 MiB that gpurun brings back: run the last two caps separately)
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench.csv
