#!/bin/bash
# Round 2, GPU call K (one B200): whole GPU suite (with the new river / plane / 4 Mi-vs-oracle tests), bench configs 1-3,
# onesweep tile size A/B.
set -u
mkdir -p gpurun_out
TAG=${1:-r02k}
timeout 2400 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
CLSPH_SORT_ITEMS=16 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${TAG}_bench_cfg2_items16.json 2> gpurun_out/${TAG}_bench_cfg2_items16.err
timeout 600 python bench.py --config config1_box_100k --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
timeout 600 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 2 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
CLSPH_SORT_ITEMS=8 timeout 600 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_cfg3_items8.json 2> gpurun_out/${TAG}_bench_cfg3_items8.err
ls -la gpurun_out | grep ${TAG}
