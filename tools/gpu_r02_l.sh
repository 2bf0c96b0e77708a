#!/bin/bash
# Round 2, GPU call L (TWO B200s): whole GPU suite (new tests: river / plane / 4 Mi vs oracle / frame export / multi-GPU at world 2),
# bench config 2 at N=1 (k_rank on the side stream, overflow pass skipped) and N=2 (exchange in two kernels).
set -u
mkdir -p gpurun_out
TAG=${1:-r02l}
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --config config1_box_100k --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
CLSPH_DIST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 50 --warmup 10 \
    > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
ls -la gpurun_out | grep ${TAG}
