#!/bin/bash
# Round 2, GPU call H (one B200): pair density with two-at-a-time list stores -- bench configs 1-3, ncu of the density kernels
# (pairs and per-particle) to compare L1 wavefronts.
set -u
mkdir -p gpurun_out
TAG=${1:-r02h}
timeout 600 python -m pytest tests -m gpu -q -x -k "pair_density or organisations or crowded" > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
timeout 600 python bench.py --config config1_box_100k --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
timeout 600 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 2 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_pairs' \
    -s 2 -c 1 -f -o gpurun_out/${TAG}_cfg2_pairs python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 \
    > gpurun_out/${TAG}_ncu_pairs.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_sub' \
    -s 2 -c 1 -f -o gpurun_out/${TAG}_cfg2_sub python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 --option pair_density=0 \
    > gpurun_out/${TAG}_ncu_sub.log 2>&1
ls -la gpurun_out | grep ${TAG}
