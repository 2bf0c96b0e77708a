#!/usr/bin/env bash
# full GPU test-suite + default bench + configs 3/4 stage times
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for cfg in config2_dambreak_1m config3_mucus_labyrinth_4m config4_river_16m; do
  timeout 900 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 0 > gpurun_out/check_$cfg.json 2> gpurun_out/check_$cfg.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/check_$cfg.json"))
    print("%-28s value %.4g p-s/s  ms/step %.3f  stage %s" % ("$cfg", d["value"], d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}))
except Exception as e:
    print("$cfg failed", e); print(open("gpurun_out/check_$cfg.err").read()[-800:])
PY
done
