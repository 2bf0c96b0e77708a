#!/usr/bin/env bash
# ncu --set full of selected kernels of the default bench; usage: gpu_prof.sh <regex> <outname> [bench args...]
set -u
mkdir -p gpurun_out
REGEX="$1"; OUT="$2"; shift 2
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s 6 -c 2 -o gpurun_out/$OUT -f \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 "$@" > gpurun_out/$OUT.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/$OUT.log | cut -c1-300
