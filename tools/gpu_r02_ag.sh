#!/bin/bash
# Round 2, GPU call AG (two B200s): final check of the committed tree -- GPU suite (multi-GPU tests included), smoke(),
# bench.py with its defaults, the reference arm, and the 2-GPU bench line.
set -u
mkdir -p gpurun_out
TAG=${1:-r02ag}
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
ls -la gpurun_out | grep ${TAG}
