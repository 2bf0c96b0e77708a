#!/bin/bash
# Round 2, GPU call R (one B200): the scaling sweep (config 5) on one GPU: uniform block over plane.obj, 1 Mi ... 64 Mi particles.
set -u
mkdir -p gpurun_out
TAG=${1:-r02r}
for c in sweep_1m sweep_4m sweep_16m; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
done
timeout 900 python bench.py --config sweep_64m --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --repeats 1 > gpurun_out/${TAG}_bench_sweep_64m.json 2> gpurun_out/${TAG}_bench_sweep_64m.err
timeout 600 python bench.py --config config4_river_16m --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --repeats 1 > gpurun_out/${TAG}_bench_cfg4_n1.json 2> gpurun_out/${TAG}_bench_cfg4_n1.err
ls -la gpurun_out | grep ${TAG}
