#!/bin/bash
# Round 2, GPU call A (one B200): ncu evidence for the kernels the bench actually times.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02_a.sh'
set -u
mkdir -p gpurun_out
OPTS="--option sub_cell_order=1 --option face_grid=1 --option fast_pairs=1 --option merged_rows=1 --option forces_blocks=4"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r02a_env.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
timeout 600 python bench.py $OPTS --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02a_bench_cfg2.json 2> gpurun_out/r02a_bench_cfg2.err
timeout 600 python bench.py $OPTS --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 2 --no-cpu-baseline \
    > gpurun_out/r02a_bench_cfg3.json 2> gpurun_out/r02a_bench_cfg3.err
timeout 600 python bench.py $OPTS --config config1_box_100k --steps 100 --warmup 10 --e2e-steps 2 --no-cpu-baseline \
    > gpurun_out/r02a_bench_cfg1.json 2> gpurun_out/r02a_bench_cfg1.err
# launch list (shares of the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_cfg2.csv \
    python bench.py $OPTS --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02a_under_ncu.log 2>&1
# full captures: neighbour/reorder/integrate kernels, then the sort kernels, config 2 and config 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_sub|k_forces_lists|k_rank|k_reorder_sub|k_integrate' \
    -s 5 -c 5 -f -o gpurun_out/r02a_cfg2_main python bench.py $OPTS --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r02a_ncu_cfg2_main.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_keys_hist|k_scan_hist|k_onesweep|k_grid_setup|k_clear_sub' \
    -s 8 -c 8 -f -o gpurun_out/r02a_cfg2_sort python bench.py $OPTS --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r02a_ncu_cfg2_sort.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_sub|k_forces_lists|k_rank|k_reorder_sub|k_integrate' \
    -s 5 -c 5 -f -o gpurun_out/r02a_cfg3_main python bench.py $OPTS --config config3_mucus_labyrinth_4m --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r02a_ncu_cfg3_main.log 2>&1
ls -la gpurun_out | tail -20
