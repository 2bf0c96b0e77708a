#!/bin/bash
# Round 2, GPU call U (TWO B200s): select-ahead with CTA-level appends vs the select kernel at N=2; smoke(); bitwise worker.
set -u
mkdir -p gpurun_out
TAG=${1:-r02u}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_smoke.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/dist_worker.py 200000 8 \
    > gpurun_out/${TAG}_dist_worker_200k.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_dist_worker_200k.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 50 --warmup 10 --e2e-steps 0 \
    > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
CLSPH_DIST_SELECT_AHEAD=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 50 --warmup 10 --e2e-steps 0 \
    > gpurun_out/${TAG}_bench_n2_selectkernel.json 2> gpurun_out/${TAG}_bench_n2_selectkernel.err
ls -la gpurun_out | grep ${TAG}
