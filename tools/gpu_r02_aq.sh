#!/bin/bash
# Round 2, GPU call AQ (one B200): ncu launch list (durations only) of bench.py's sub-steps on the final tree.
set -u
mkdir -p gpurun_out
TAG=${1:-r02aq}
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/${TAG}_launches_config2.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-large-point > gpurun_out/${TAG}_config2.log 2>&1
ls -la gpurun_out | grep ${TAG}
