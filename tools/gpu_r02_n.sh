#!/bin/bash
# Round 2, GPU call N (one B200): defaults after the row prefetch and the factored forces on self-listed lists; A/B against
# factored_forces=0; ncu capture of one whole sub-step of the defaults (configs 2 and 3).
set -u
mkdir -p gpurun_out
TAG=${1:-r02n}
for c in config2_dambreak_1m config3_mucus_labyrinth_4m config1_box_100k; do
  timeout 300 python bench.py --config $c --steps 30 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 2 > gpurun_out/${TAG}_${c}_default.json 2> gpurun_out/${TAG}_${c}_default.err
  timeout 300 python bench.py --config $c --steps 30 --warmup 10 --no-cpu-baseline --e2e-steps 0 --repeats 2 --option factored_forces=0 > gpurun_out/${TAG}_${c}_plainforces.json 2> gpurun_out/${TAG}_${c}_plainforces.err
done
timeout 600 python -m pytest tests -m gpu -q -x -k "organisations or crowded or developed or million or golden or pair" > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_pairs|k_forces_lists|k_reorder_sub|k_rank|k_integrate|k_keys_hist|k_onesweep|k_clear_sub|k_grid_setup|k_scan_hist|k_forces_sub' \
    -s 20 -c 14 -f -o gpurun_out/${TAG}_cfg2_substep python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 \
    > gpurun_out/${TAG}_ncu_cfg2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_pairs|k_forces_lists|k_reorder_sub|k_rank|k_integrate' \
    -s 6 -c 5 -f -o gpurun_out/${TAG}_cfg3_main python bench.py --config config3_mucus_labyrinth_4m --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 \
    > gpurun_out/${TAG}_ncu_cfg3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --repeats 0 > gpurun_out/${TAG}_launches.log 2>&1
ls -la gpurun_out | grep ${TAG}
