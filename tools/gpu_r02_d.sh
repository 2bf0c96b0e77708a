#!/bin/bash
# Round 2, GPU call D (one B200): tile density with the packed-half pre-test + neighbour lists, list force pass with the tile pair terms.
set -u
mkdir -p gpurun_out
TAG=${1:-r02d}
T="sub_cell_order=1,face_grid=1,fast_pairs=1,tile_kernels=1,forces_blocks=4"
OPTS="--option sub_cell_order=1 --option face_grid=1 --option fast_pairs=1 --option tile_kernels=1 --option forces_blocks=4"
CLSPH_OPTIONS="$T" timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_tiles.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_tiles.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m libclsph_b200.selfcheck --config config3_mucus_labyrinth_4m \
      --particles 30000 --timed-steps 2 --set $T > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 600 python bench.py $OPTS --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
timeout 600 python bench.py $OPTS --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 2 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_tiles|k_density_slow|k_forces_lists_tile|k_forces_sub' \
    -s 4 -c 4 -f -o gpurun_out/${TAG}_cfg2_tiles python bench.py $OPTS --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/${TAG}_ncu_cfg2.log 2>&1
ls -la gpurun_out | grep ${TAG}
