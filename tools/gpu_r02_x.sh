#!/bin/bash
# Round 2, GPU call X (one B200): direct force kernel with the list prefetch.
set -u
mkdir -p gpurun_out
TAG=${1:-r02x}
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_cfg2.json 2> gpurun_out/${TAG}_cfg2.err
timeout 300 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_cfg3.json 2> gpurun_out/${TAG}_cfg3.err
timeout 300 python bench.py --config sweep_16m --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --repeats 1 > gpurun_out/${TAG}_sweep16m.json 2> gpurun_out/${TAG}_sweep16m.err
timeout 600 python -m pytest tests -m gpu -q -x -k "organisations or crowded or developed or million or golden" > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
