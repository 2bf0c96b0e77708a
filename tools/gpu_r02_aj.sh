#!/bin/bash
# Round 2, GPU call AJ (one B200): per-launch durations (ncu, cold and serialised) of the sub-step with the counting sort.
set -u
mkdir -p gpurun_out
TAG=${1:-r02aj}
for cfg in config2_dambreak_1m config1_box_100k; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches_${cfg}.csv \
    python bench.py --config $cfg --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-large-point > gpurun_out/${TAG}_${cfg}.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
