#!/bin/bash
# Round 2, GPU call AM (eight B200s): the counting sort across GPUs -- bitwise worker at world 8, bench lines at 8 GPUs
# (config 2 per GPU with its parity record, config 4 river, the 64 Mi block), and at 2 and 4 GPUs.
set -u
mkdir -p gpurun_out
TAG=${1:-r02am}
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $RUN --nproc-per-node 8 --master-port 29801 tests/dist_worker.py 400000 8 > gpurun_out/${TAG}_dist_worker_w8.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_dist_worker_w8.log
timeout 600 $RUN --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 --steps 50 --warmup 10 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err
timeout 600 $RUN --nproc-per-node 8 --master-port 29803 bench.py --gpus 8 --steps 50 --warmup 10 --option count_sort=0 --e2e-steps 0 > gpurun_out/${TAG}_bench_n8_radix.json 2> gpurun_out/${TAG}_bench_n8_radix.err
timeout 600 $RUN --nproc-per-node 8 --master-port 29804 bench.py --gpus 8 --config config4_river_16m --steps 30 --warmup 5 --e2e-steps 0 > gpurun_out/${TAG}_bench_n8_cfg4.json 2> gpurun_out/${TAG}_bench_n8_cfg4.err
timeout 600 $RUN --nproc-per-node 4 --master-port 29805 bench.py --gpus 4 --steps 50 --warmup 10 --e2e-steps 0 > gpurun_out/${TAG}_bench_n4.json 2> gpurun_out/${TAG}_bench_n4.err
timeout 600 $RUN --nproc-per-node 2 --master-port 29806 bench.py --gpus 2 --steps 50 --warmup 10 --e2e-steps 0 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
ls -la gpurun_out | grep ${TAG}
