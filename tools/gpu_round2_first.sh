#!/bin/bash
# First GPU call of round 2 (one B200): the paths written on the CPU emulator at the end of round 1
# (sub-cell order, face grid) meet a GPU for the first time. Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
set -u
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv; } > gpurun_out/r02_env.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
# the established parity suite, unchanged, on the candidate kernels (CLSPH_OPTIONS applies to every context)
CLSPH_OPTIONS="sub_cell_order=1,face_grid=1,fast_pairs=1,merged_rows=1,forces_blocks=4" timeout 1200 python -m pytest tests/test_gpu_parity.py \
    -m gpu -q > gpurun_out/r02_pytest_gpu_parity_on_candidate.log 2>&1
for cfg in config2_dambreak_1m config3_mucus_labyrinth_4m; do
  timeout 600 python -m libclsph_b200.selfcheck --config $cfg --set sub_cell_order=1,face_grid=1,fast_pairs=1 --set sub_cell_order=1,face_grid=1,fast_pairs=1,merged_rows=1 --set sub_cell_order=1,face_grid=1,fast_pairs=1,merged_rows=1,forces_blocks=4 --set sub_cell_order=1,face_grid=1,fast_pairs=1,deferred_lists=1,forces_blocks=4 --set sub_cell_order=1,face_grid=1,fast_pairs=1,forces_blocks=4 --set face_grid=1,fast_pairs=1,forces_blocks=4 \
      > gpurun_out/r02_selfcheck_$cfg.json 2> gpurun_out/r02_selfcheck_$cfg.err
done
# memcheck + racecheck of the new kernels on a small workload (established and candidate sets in one run)
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m libclsph_b200.selfcheck --config config3_mucus_labyrinth_4m \
      --particles 30000 --timed-steps 2 --set sub_cell_order=1,face_grid=1,fast_pairs=1 --set sub_cell_order=1,face_grid=1,fast_pairs=1,deferred_lists=1,forces_blocks=4 --set face_grid=1,fast_pairs=1,forces_blocks=4 \
      > gpurun_out/r02_sanitizer_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitizer_$tool.log
done
for org in default candidate; do
  timeout 600 python bench.py --organisation $org --steps 50 --warmup 10 > gpurun_out/r02_bench_cfg2_$org.json 2> gpurun_out/r02_bench_cfg2_$org.err
  timeout 600 python bench.py --organisation $org --config config3_mucus_labyrinth_4m --steps 10 --warmup 3 --e2e-steps 2 --no-cpu-baseline \
      > gpurun_out/r02_bench_cfg3_$org.json 2> gpurun_out/r02_bench_cfg3_$org.err
done
timeout 600 python bench.py > gpurun_out/r02_bench_auto.json 2> gpurun_out/r02_bench_auto.err
# launch list (shares) and one full capture of the new kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_candidate.csv \
    python bench.py --organisation candidate --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_sub|k_forces_lists|k_rank|k_reorder_sub|k_integrate' \
    -s 10 -c 5 -o gpurun_out/r02_candidate_kernels python bench.py --organisation candidate --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
