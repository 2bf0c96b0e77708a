#!/bin/bash
# Round 2, GPU call F (one B200): k_density_pairs (default) -- GPU suite, bench configs 1-3, ncu capture of one sub-step.
set -u
mkdir -p gpurun_out
TAG=${1:-r02f}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --option pair_density=0 --e2e-steps 0 > gpurun_out/${TAG}_bench_cfg2_nopairs.json 2> gpurun_out/${TAG}_bench_cfg2_nopairs.err
timeout 600 python bench.py --config config1_box_100k --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
timeout 600 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 2 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_pairs|k_forces_lists|k_reorder_sub|k_rank|k_integrate' \
    -s 10 -c 5 -f -o gpurun_out/${TAG}_cfg2_main python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 \
    > gpurun_out/${TAG}_ncu_cfg2.log 2>&1
ls -la gpurun_out | grep ${TAG}
