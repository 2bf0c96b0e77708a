#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02y}
for d in 1 3; do
CLSPH_FORCES_DIRECT=$d timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_cfg2_d$d.json 2> gpurun_out/${TAG}_cfg2_d$d.err
CLSPH_FORCES_DIRECT=$d timeout 300 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_cfg3_d$d.json 2> gpurun_out/${TAG}_cfg3_d$d.err
done
