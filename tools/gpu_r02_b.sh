#!/bin/bash
# Round 2, GPU call B (one B200): the tile kernels (tiles.cu) meet a GPU for the first time.
set -u
mkdir -p gpurun_out
T="sub_cell_order=1,face_grid=1,fast_pairs=1,tile_kernels=1"
OPTS="--option sub_cell_order=1 --option face_grid=1 --option fast_pairs=1 --option tile_kernels=1"
for cfg in config2_dambreak_1m config3_mucus_labyrinth_4m; do
  timeout 600 python -m libclsph_b200.selfcheck --config $cfg --set sub_cell_order=1,face_grid=1,fast_pairs=1,merged_rows=1,forces_blocks=4 --set $T \
      > gpurun_out/r02b_selfcheck_$cfg.json 2> gpurun_out/r02b_selfcheck_$cfg.err
done
# the whole GPU parity suite on the tile kernels (CLSPH_OPTIONS applies to every context)
CLSPH_OPTIONS="$T" timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02b_pytest_tiles.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_pytest_tiles.log
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m libclsph_b200.selfcheck --config config3_mucus_labyrinth_4m \
      --particles 30000 --timed-steps 2 --set $T > gpurun_out/r02b_sanitizer_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_sanitizer_$tool.log
done
timeout 600 python bench.py $OPTS --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02b_bench_cfg2.json 2> gpurun_out/r02b_bench_cfg2.err
timeout 600 python bench.py $OPTS --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 2 --no-cpu-baseline \
    > gpurun_out/r02b_bench_cfg3.json 2> gpurun_out/r02b_bench_cfg3.err
timeout 600 python bench.py $OPTS --config config1_box_100k --steps 100 --warmup 10 --e2e-steps 2 --no-cpu-baseline \
    > gpurun_out/r02b_bench_cfg1.json 2> gpurun_out/r02b_bench_cfg1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_tiles|k_forces_tiles|k_density_slow|k_forces_slow' \
    -s 4 -c 4 -f -o gpurun_out/r02b_cfg2_tiles python bench.py $OPTS --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r02b_ncu_cfg2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_tiles|k_forces_tiles|k_density_slow|k_forces_slow' \
    -s 4 -c 4 -f -o gpurun_out/r02b_cfg3_tiles python bench.py $OPTS --config config3_mucus_labyrinth_4m --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r02b_ncu_cfg3.log 2>&1
ls -la gpurun_out | grep r02b
