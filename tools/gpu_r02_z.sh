#!/bin/bash
# Round 2, GPU call Z (one B200): resident waves with SM-local ranges (forces, density) on/off.
set -u
mkdir -p gpurun_out
TAG=${1:-r02z}
for f in 1 0; do for d in 1 0; do
CLSPH_FORCES_PERSIST=$f CLSPH_DENSITY_PERSIST=$d timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_cfg2_f${f}d${d}.json 2> gpurun_out/${TAG}_cfg2_f${f}d${d}.err
CLSPH_FORCES_PERSIST=$f CLSPH_DENSITY_PERSIST=$d timeout 300 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_cfg3_f${f}d${d}.json 2> gpurun_out/${TAG}_cfg3_f${f}d${d}.err
done; done
timeout 600 ncu --set full --clock-control none -k regex:'k_forces_lists_direct|k_density_pairs' -s 4 -c 2 -f -o gpurun_out/${TAG}_waves python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 > gpurun_out/${TAG}_ncu.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "organisations or crowded or developed or million or golden or pair" > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
