#!/bin/bash
# Round 2, GPU call E (one B200): the flipped library defaults under the whole GPU suite, bench lines of configs 1-3 with the
# defaults, pipe-rate microbenchmark (FFMA2 / HFMA2 issue rates).
set -u
mkdir -p gpurun_out
TAG=${1:-r02e}
./build/pipe_rates > gpurun_out/${TAG}_pipe_rates.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
timeout 600 python bench.py --config config1_box_100k --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
timeout 600 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 2 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
ls -la gpurun_out | grep ${TAG}
