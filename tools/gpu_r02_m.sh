#!/bin/bash
# Round 2, GPU call M (one B200): force pass A/B -- factored pair terms on/off x resident CTAs 3/4 x neighbours per trip 2/4.
set -u
mkdir -p gpurun_out
TAG=${1:-r02m}
run() {  # name, env, options...
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 300 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 0 "$@" > gpurun_out/${TAG}_cfg2_$name.json 2> gpurun_out/${TAG}_cfg2_$name.err
  env $envs timeout 300 python bench.py --config config3_mucus_labyrinth_4m --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 0 --repeats 0 "$@" > gpurun_out/${TAG}_cfg3_$name.json 2> gpurun_out/${TAG}_cfg3_$name.err
}
run plain_b4 X=1 --option factored_forces=0
run fact_b4_t2 CLSPH_FORCES_TRIP=2
run fact_b4_t4 CLSPH_FORCES_TRIP=4
run fact_b3_t2 CLSPH_FORCES_TRIP=2 --option forces_blocks=3
run fact_b3_t4 CLSPH_FORCES_TRIP=4 --option forces_blocks=3
timeout 600 python -m pytest tests -m gpu -q -x -k "organisations or crowded or developed or million or golden" > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_forces_lists_tile' \
    -s 2 -c 1 -f -o gpurun_out/${TAG}_cfg2_forces python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 \
    > gpurun_out/${TAG}_ncu_forces.log 2>&1
ls -la gpurun_out | grep ${TAG}
