#!/bin/bash
# Round 2, GPU call P (TWO B200s): exchange in place -- multi-GPU tests, bitwise worker, bench at N=2 in place vs copying.
set -u
mkdir -p gpurun_out
TAG=${1:-r02p}
timeout 900 python -m pytest tests -m gpu -q -x -k "slab or multi" > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/dist_worker.py 200000 8 \
    > gpurun_out/${TAG}_dist_worker_200k.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_dist_worker_200k.log
CLSPH_DIST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 50 --warmup 10 --e2e-steps 5 \
    > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
CLSPH_DIST_IN_PLACE=0 CLSPH_DIST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 50 --warmup 10 --e2e-steps 0 \
    > gpurun_out/${TAG}_bench_n2_copy.json 2> gpurun_out/${TAG}_bench_n2_copy.err
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
ls -la gpurun_out | grep ${TAG}
