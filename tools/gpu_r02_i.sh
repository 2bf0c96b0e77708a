#!/bin/bash
# Round 2, GPU call I (one B200): sweep of the pair-density variants (walk 0/1/2 x entries stored one / two at a time).
set -u
mkdir -p gpurun_out
TAG=${1:-r02i}
for v in 0 1 2 3 4 5; do
  timeout 300 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 0 --option pair_variant=$v \
      > gpurun_out/${TAG}_cfg2_v$v.json 2> gpurun_out/${TAG}_cfg2_v$v.err
  timeout 300 python bench.py --config config3_mucus_labyrinth_4m --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 0 --repeats 0 --option pair_variant=$v \
      > gpurun_out/${TAG}_cfg3_v$v.json 2> gpurun_out/${TAG}_cfg3_v$v.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_pairs' \
    -s 2 -c 1 -f -o gpurun_out/${TAG}_cfg2_pairs_v1 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 --option pair_variant=1 \
    > gpurun_out/${TAG}_ncu_pairs.log 2>&1
ls -la gpurun_out | grep ${TAG}
