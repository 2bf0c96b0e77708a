#!/bin/bash
# Round 2, GPU call S (one B200): force pass without the shared-memory tile (CLSPH_FORCES_DIRECT) against the default.
set -u
mkdir -p gpurun_out
TAG=${1:-r02s}
run() {  # name, env, options...
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 300 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 1 "$@" > gpurun_out/${TAG}_cfg2_$name.json 2> gpurun_out/${TAG}_cfg2_$name.err
  env $envs timeout 300 python bench.py --config config3_mucus_labyrinth_4m --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 0 --repeats 1 "$@" > gpurun_out/${TAG}_cfg3_$name.json 2> gpurun_out/${TAG}_cfg3_$name.err
}
run default X=1
run direct1_plain CLSPH_FORCES_DIRECT=1 --option factored_forces=0
run direct1_fact CLSPH_FORCES_DIRECT=1 --option factored_forces=1
run direct2_plain CLSPH_FORCES_DIRECT=2 --option factored_forces=0
run direct2_fact CLSPH_FORCES_DIRECT=2 --option factored_forces=1
CLSPH_FORCES_DIRECT=1 timeout 600 python -m pytest tests -m gpu -q -x -k "organisations or crowded or developed or million or golden" > gpurun_out/${TAG}_pytest_direct.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_direct.log
CLSPH_FORCES_DIRECT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_forces_lists_direct' \
    -s 2 -c 1 -f -o gpurun_out/${TAG}_cfg2_forces_direct python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 --option factored_forces=1 \
    > gpurun_out/${TAG}_ncu_forces.log 2>&1
ls -la gpurun_out | grep ${TAG}
