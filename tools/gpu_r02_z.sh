#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02z}
for d in 1 0; do
CLSPH_FORCES_PERSIST=$d timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_cfg2_p$d.json 2> gpurun_out/${TAG}_cfg2_p$d.err
CLSPH_FORCES_PERSIST=$d timeout 300 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_cfg3_p$d.json 2> gpurun_out/${TAG}_cfg3_p$d.err
done
timeout 600 ncu --set full --clock-control none -k regex:'k_forces_lists_direct' -s 2 -c 1 -f -o gpurun_out/${TAG}_forces python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 > gpurun_out/${TAG}_ncu.log 2>&1
