#!/bin/bash
# Round 2, GPU call C (one B200): tile kernels after the fixes listed in profiles/r02_b_summary.md.
set -u
mkdir -p gpurun_out
T="sub_cell_order=1,face_grid=1,fast_pairs=1,tile_kernels=1"
OPTS="--option sub_cell_order=1 --option face_grid=1 --option fast_pairs=1 --option tile_kernels=1"
CLSPH_OPTIONS="$T" timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest_tiles.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_pytest_tiles.log
timeout 600 python bench.py $OPTS --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02c_bench_cfg2.json 2> gpurun_out/r02c_bench_cfg2.err
timeout 600 python bench.py $OPTS --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 2 --no-cpu-baseline \
    > gpurun_out/r02c_bench_cfg3.json 2> gpurun_out/r02c_bench_cfg3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_tiles|k_forces_tiles|k_density_slow|k_forces_slow' \
    -s 4 -c 4 -f -o gpurun_out/r02c_cfg2_tiles python bench.py $OPTS --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r02c_ncu_cfg2.log 2>&1
ls -la gpurun_out | grep r02c
