#!/bin/bash
# Round 2, GPU call W (one B200): what the driver runs at round end -- GPU suite, smoke(), bench.py with its defaults, the reference arm.
set -u
mkdir -p gpurun_out
TAG=${1:-r02w}
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 600 python bench.py --config config3_mucus_labyrinth_4m --steps 20 --warmup 5 --e2e-steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
timeout 600 python bench.py --config config1_box_100k --steps 100 --warmup 10 --e2e-steps 20 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
ls -la gpurun_out | grep ${TAG}
