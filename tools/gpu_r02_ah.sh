#!/bin/bash
# Round 2, GPU call AH (one B200): counting sort on the sub-cell table against the radix passes -- parity tests, memcheck,
# and the A/B of the bench lines (configs 2 and 3, both orders so that drift shows).
set -u
mkdir -p gpurun_out
TAG=${1:-r02ah}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_new_paths.py -q -x -k "counting or organisations or resident or binary_search or one_million or four_million or bitwise" > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "counting and mucus" > gpurun_out/${TAG}_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_memcheck.log
for rep in 1 2; do
  for cs in 1 0; do
    timeout 600 python bench.py --steps 100 --warmup 10 --e2e-steps 3 --no-cpu-baseline --no-large-point --option count_sort=$cs > gpurun_out/${TAG}_cfg2_cs${cs}_${rep}.json 2> gpurun_out/${TAG}_cfg2_cs${cs}_${rep}.err
  done
done
for cs in 1 0; do
  timeout 600 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 3 --no-cpu-baseline --option count_sort=$cs > gpurun_out/${TAG}_cfg3_cs${cs}.json 2> gpurun_out/${TAG}_cfg3_cs${cs}.err
  timeout 600 python bench.py --config config1_box_100k --steps 100 --warmup 10 --e2e-steps 3 --no-cpu-baseline --option count_sort=$cs > gpurun_out/${TAG}_cfg1_cs${cs}.json 2> gpurun_out/${TAG}_cfg1_cs${cs}.err
done
ls -la gpurun_out | grep ${TAG}
