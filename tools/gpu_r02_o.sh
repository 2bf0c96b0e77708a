#!/bin/bash
# Round 2, GPU call O (one B200): end-to-end A/B of the zero-copy upload; whole GPU suite; bench configs 1-3 with the folded launches.
set -u
mkdir -p gpurun_out
TAG=${1:-r02o}
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 30 > gpurun_out/${TAG}_bench_cfg2_zc1.json 2> gpurun_out/${TAG}_bench_cfg2_zc1.err
CLSPH_ZERO_COPY_UPLOAD=0 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 30 > gpurun_out/${TAG}_bench_cfg2_zc0.json 2> gpurun_out/${TAG}_bench_cfg2_zc0.err
timeout 600 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_zc1.json 2> gpurun_out/${TAG}_bench_cfg3_zc1.err
CLSPH_ZERO_COPY_UPLOAD=0 timeout 600 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_zc0.json 2> gpurun_out/${TAG}_bench_cfg3_zc0.err
timeout 600 python bench.py --config config1_box_100k --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 30 > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
ls -la gpurun_out | grep ${TAG}
