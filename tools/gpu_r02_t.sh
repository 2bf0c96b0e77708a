#!/bin/bash
# Round 2, GPU call T (TWO B200s): the integrator prepares the exchange -- multi-GPU tests, bitwise worker, bench N=2 (vs select kernel), N=1.
set -u
mkdir -p gpurun_out
TAG=${1:-r02t}
timeout 900 python -m pytest tests -m gpu -q -x -k "slab or multi" > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/dist_worker.py 200000 8 \
    > gpurun_out/${TAG}_dist_worker_200k.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_dist_worker_200k.log
CLSPH_DIST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 50 --warmup 10 --e2e-steps 5 \
    > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
CLSPH_DIST_SELECT_AHEAD=0 CLSPH_DIST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 50 --warmup 10 --e2e-steps 0 \
    > gpurun_out/${TAG}_bench_n2_selectkernel.json 2> gpurun_out/${TAG}_bench_n2_selectkernel.err
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
ls -la gpurun_out | grep ${TAG}
