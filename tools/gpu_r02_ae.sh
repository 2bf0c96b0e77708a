#!/bin/bash
# Round 2, GPU call AE (one B200): pair density with the hits staged in shared memory (pair_variant=6) against the default (5).
set -u
mkdir -p gpurun_out
TAG=${1:-r02ae}
for v in 5 6; do
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 2 --option pair_variant=$v > gpurun_out/${TAG}_cfg2_v$v.json 2> gpurun_out/${TAG}_cfg2_v$v.err
timeout 300 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --repeats 2 --option pair_variant=$v > gpurun_out/${TAG}_cfg3_v$v.json 2> gpurun_out/${TAG}_cfg3_v$v.err
timeout 300 python bench.py --config sweep_16m --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --repeats 1 --option pair_variant=$v > gpurun_out/${TAG}_sweep16m_v$v.json 2> gpurun_out/${TAG}_sweep16m_v$v.err
done
timeout 600 python -m pytest tests -m gpu -q -x -k "organisations or crowded or pair_density" > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_density_pairs' -s 2 -c 1 -f -o gpurun_out/${TAG}_staged python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 --option pair_variant=6 > gpurun_out/${TAG}_ncu.log 2>&1
