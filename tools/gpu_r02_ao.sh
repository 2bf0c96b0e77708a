#!/bin/bash
# Round 2, GPU call AO (one B200, the last minutes of the budget): the driver's sequence on the final tree.
set -u
mkdir -p gpurun_out
TAG=${1:-r02ao}
timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_smoke.log
timeout 60 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
timeout 30 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
timeout 20 python bench.py --config config1_box_100k --steps 100 --warmup 10 --e2e-steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
ls -la gpurun_out | grep ${TAG}
