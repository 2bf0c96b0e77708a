#!/bin/bash
# Round 2, GPU call AN (eight B200s): after the row-wise zeroing of the sub-cell table -- bitwise worker, bench at 8 GPUs.
set -u
mkdir -p gpurun_out
TAG=${1:-r02an}
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $RUN --nproc-per-node 8 --master-port 29811 tests/dist_worker.py 200000 6 > gpurun_out/${TAG}_dist_worker_w8.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_dist_worker_w8.log
timeout 600 $RUN --nproc-per-node 8 --master-port 29812 bench.py --gpus 8 --steps 50 --warmup 10 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err
timeout 600 $RUN --nproc-per-node 8 --master-port 29813 bench.py --gpus 8 --config config4_river_16m --steps 30 --warmup 5 --e2e-steps 0 > gpurun_out/${TAG}_bench_n8_cfg4.json 2> gpurun_out/${TAG}_bench_n8_cfg4.err
ls -la gpurun_out | grep ${TAG}
