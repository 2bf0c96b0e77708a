#!/bin/bash
# Round 2, GPU call Q (EIGHT B200s): the exchange in place + device-scope fences -- bitwise check at world 8, config 2 weak scaling at
# 8 and 4 GPUs, config 4 (river.obj, 16 Mi over 8 GPUs), the 64 Mi and 16 Mi points of the sweep over 8 GPUs.
set -u
mkdir -p gpurun_out
TAG=${1:-r02q}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29721 tests/dist_worker.py 480000 6 > gpurun_out/${TAG}_dist_worker_w8.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_dist_worker_w8.log
CLSPH_DIST_TIMING=1 timeout 400 $TR --nproc-per-node 8 --master-port 29722 bench.py --gpus 8 --steps 50 --warmup 10 \
    > gpurun_out/${TAG}_bench_cfg2_n8.json 2> gpurun_out/${TAG}_bench_cfg2_n8.err
timeout 400 $TR --nproc-per-node 4 --master-port 29723 bench.py --gpus 4 --steps 50 --warmup 10 --e2e-steps 5 \
    > gpurun_out/${TAG}_bench_cfg2_n4.json 2> gpurun_out/${TAG}_bench_cfg2_n4.err
timeout 500 $TR --nproc-per-node 8 --master-port 29725 bench.py --gpus 8 --config config4_river_16m --steps 30 --warmup 10 --e2e-steps 3 --no-parity \
    > gpurun_out/${TAG}_bench_cfg4_n8.json 2> gpurun_out/${TAG}_bench_cfg4_n8.err
timeout 500 $TR --nproc-per-node 8 --master-port 29726 bench.py --gpus 8 --config sweep_64m --particles 8388608 --steps 20 --warmup 5 --e2e-steps 0 --no-parity \
    > gpurun_out/${TAG}_bench_sweep64m_n8.json 2> gpurun_out/${TAG}_bench_sweep64m_n8.err
timeout 500 $TR --nproc-per-node 8 --master-port 29727 bench.py --gpus 8 --config sweep_16m --particles 2097152 --steps 30 --warmup 5 --e2e-steps 0 --no-parity \
    > gpurun_out/${TAG}_bench_sweep16m_n8.json 2> gpurun_out/${TAG}_bench_sweep16m_n8.err
ls -la gpurun_out | grep ${TAG}
