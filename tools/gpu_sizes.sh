#!/usr/bin/env bash
# throughput across sizes and configs (device-resident number + stage times), one JSON per run
set -u
mkdir -p gpurun_out
for cfg in config1_box_100k sweep_1m sweep_4m sweep_16m sweep_64m config3_mucus_labyrinth_4m config4_river_16m; do
  steps=30; [ "$cfg" = sweep_64m ] && steps=10
  timeout 900 python bench.py --config $cfg --steps $steps --warmup 5 --no-cpu-baseline --e2e-steps 0 > gpurun_out/sizes_$cfg.json 2> gpurun_out/sizes_$cfg.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sizes_$cfg.json"))
    print("%-28s n=%9d value %.4g p-s/s  ms/step %.3f  step-frac %.4f  stage %s" % ("$cfg", d["config"]["particles_per_gpu"], d["value"], d["ms_per_step"], d["roofline"]["whole_step"]["frac"], {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}))
except Exception as e:
    print("$cfg failed", e); print(open("gpurun_out/sizes_$cfg.err").read()[-800:])
PY
done
