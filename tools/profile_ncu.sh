#!/bin/bash
# ncu recipe behind profiles/ (run on a B200 box, e.g. `gpurun -- bash tools/profile_ncu.sh`).
# Successor of the reference's trace tooling (event0/event1 markers + scripts/parse_trace.py +
# profile/plot_*.py): launch list (per-launch device time; compare SHARES, the numbers are
# cold-cache and serialised) and one `--set full` capture of the dominant kernel.
# Outputs land in gpurun_out/; summarise with:
#   ncu -i gpurun_out/prof_fused.ncu-rep --page raw --csv   (dram bytes, pipe utilisation, stalls)
#   ncu -i gpurun_out/prof_fused.ncu-rep --page source --csv (per-instruction stall samples)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
# a command is profiled only after the same command exited 0 without ncu
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv $CMD > /dev/null 2>&1
CMD2="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/plain2.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:fused_gs4096 -s 3 -c 1 -f \
    -o gpurun_out/prof_fused $CMD2 > gpurun_out/ncu_full.log 2>&1
# secondary kernels (tile / column / CT / small-N), one launch each
CMD3="python tools/bench_configs.py --reps 3 --configs ntt,3,4"
$CMD3 > gpurun_out/plain3.log 2>&1 || { echo "plain run failed"; exit 1; }
for k in tile_gs_kernel column_kernel tile_ct fused_gs_small_kernel poly_gs_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f \
      -o gpurun_out/prof_$k $CMD3 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
