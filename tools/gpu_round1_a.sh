#!/usr/bin/env bash
# First GPU contact: smoke, parity tests, sanitizer on the smoke step, short bench, environment probes.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
{
  echo "== env"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" ;
  echo "== opencl probe"; ls /etc/OpenCL/vendors 2>&1; find / -name 'libnvidia-opencl*' -o -name 'libpocl*' 2>/dev/null | head
} > gpurun_out/env.log 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -15 gpurun_out/sanitizer.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; cat gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
