#!/bin/bash
# Multi-GPU measurements on one box (run under `gpurun --gpus 8`):
#   batched bench (weak scaling, no communication) at 1/2/4/8 GPUs,
#   cfg4 shards (N=2^16 x 4096 polys split over G GPUs) via bench_configs on rank-local data,
#   four-step N=2^26 at 2/4/8 GPUs with whole-vector parity against the golden.
# Results land in gpurun_out/multigpu_*.json(l).
set -u
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
echo "GPUs: $G"
run_tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
: > gpurun_out/multigpu_bench.jsonl
: > gpurun_out/multigpu_fourstep.jsonl
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' >> gpurun_out/multigpu_bench.jsonl
for n in 2 4 8; do
  [ $n -le $G ] || continue
  run_tr $n bench.py --gpus $n --steps 20 --warmup 3 2>/dev/null | grep '^{' >> gpurun_out/multigpu_bench.jsonl
done
python tools/fourstep_run.py --logn 26 --verify --steps 5 2>/dev/null | grep '^{' >> gpurun_out/multigpu_fourstep.jsonl
for n in 2 4 8; do
  [ $n -le $G ] || continue
  run_tr $n tools/fourstep_run.py --logn 26 --verify --steps 5 2>/dev/null | grep '^{' >> gpurun_out/multigpu_fourstep.jsonl
done
for n in 2 4 8; do
  [ $n -le $G ] || continue
  run_tr $n tools/fourstep_run.py --logn 26 --verify --steps 5 --fused 2>/dev/null | grep '^{' >> gpurun_out/multigpu_fourstep.jsonl
done
python - <<'PY'
import json
for f in ("gpurun_out/multigpu_bench.jsonl", "gpurun_out/multigpu_fourstep.jsonl"):
    for l in open(f):
        d = json.loads(l)
        if "metric" in d:
            print("bench n_gpus", d["n_gpus"], "polys/s %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"],
                  "e2e %.4g" % (d["e2e"]["value"] or 0), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        else:
            print("fourstep", d.get("exchange", "")[:5], "n_gpus", d["n_gpus"], "ms dev-order %.3f nat %.3f" % (d["ms_transposed_order"], d["ms_natural_order"]),
                  "a2a ms", d["all_to_all_ms"], "GB/s/dir", d["all_to_all_GBps_per_gpu_per_dir"], "exact", d["bit_exact_vs_golden"])
PY
