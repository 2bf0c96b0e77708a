#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02aa}
for c in config2_dambreak_1m config3_mucus_labyrinth_4m config1_box_100k; do
timeout 300 python bench.py --config $c --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err
done
timeout 600 python -m pytest tests -m gpu -q -x -k "organisations or crowded or developed or million or golden or lattice" > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
