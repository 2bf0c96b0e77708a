#!/usr/bin/env bash
# multi-GPU checks: worker test at world 2 (and 4 if available), then single-GPU suite sanity
set -u
mkdir -p gpurun_out
nvidia-smi -L
NG=$(nvidia-smi -L | wc -l)
echo "== dist worker world=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_worker.py 60000 8 > gpurun_out/dist2.log 2>&1; echo "rc=$?"; grep -v "^W\|^\[W\|Warning" gpurun_out/dist2.log | tail -25
if [ "$NG" -ge 4 ]; then echo "== dist worker world=4"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 tests/dist_worker.py 200000 4 > gpurun_out/dist4.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/dist4.log; fi
echo "== single-GPU suite"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -k "not multi_gpu" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench --gpus 2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 50 --warmup 10 --e2e-steps 5 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "rc=$?"; tail -c 2500 gpurun_out/bench_g2.json; grep -v "^W\|Warning\|^\*\|OMP_NUM" gpurun_out/bench_g2.err | tail -5
echo "== bench --gpus 1 (same box)"; timeout 900 python bench.py --steps 50 --warmup 10 --e2e-steps 5 --no-cpu-baseline > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g1.json')); print('g1 value %.4g ms/step %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
