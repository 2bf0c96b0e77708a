#!/usr/bin/env bash
# Second GPU contact: whole GPU suite (no -x), launch list + full ncu capture of the two neighbour kernels,
# bench at the default config.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -60 gpurun_out/pytest_gpu.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (forces, density)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_forces|k_density' -s 4 -c 2 -o gpurun_out/prof_r01_neighbors -f \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu2.log 2>&1; echo "ncu full rc=$?"
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"; cat gpurun_out/bench_b.json; tail -3 gpurun_out/bench_b.err
