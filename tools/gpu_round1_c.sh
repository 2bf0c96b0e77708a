#!/usr/bin/env bash
# list mode vs two-pass mode: tests, then bench both, then bench configs 1 and 3 quickly
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
for mode in 0 1; do
  echo "== bench neighbour_lists=$mode"
  timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 5 --option neighbour_lists=$mode > gpurun_out/bench_c_mode$mode.json 2> gpurun_out/bench_c_mode$mode.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c_mode$mode.json"))
print("value %.4g ms/step %.4f stage %s" % (d["value"], d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["stage_ms"].items()}))
PY
done
echo "== bench config3 (mucus labyrinth 4M), lists"
timeout 900 python bench.py --config config3_mucus_labyrinth_4m --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_c_cfg3.json 2> gpurun_out/bench_c_cfg3.err; tail -c 1500 gpurun_out/bench_c_cfg3.json; tail -3 gpurun_out/bench_c_cfg3.err
