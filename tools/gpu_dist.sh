#!/usr/bin/env bash
# multi-GPU checks: worker test at world 2 (and 4 if available), then single-GPU suite sanity
set -u
mkdir -p gpurun_out
nvidia-smi -L
NG=$(nvidia-smi -L | wc -l)
echo "== dist worker world=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_worker.py 60000 4 > gpurun_out/dist2.log 2>&1; echo "rc=$?"; grep -v "^W\|^\[W\|Warning" gpurun_out/dist2.log | tail -25
if [ "$NG" -ge 4 ]; then echo "== dist worker world=4"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 tests/dist_worker.py 200000 4 > gpurun_out/dist4.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/dist4.log; fi
echo "== single-GPU suite"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -k "not multi_gpu" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
