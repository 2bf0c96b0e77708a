#!/usr/bin/env bash
# scaling run on up to 8 GPUs: correctness worker at the largest world, then bench at N = 1, 2, 4, 8
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
W=$NG
echo "== dist worker world=$W"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29621 tests/dist_worker.py $((W*60000)) 5 > gpurun_out/dist$W.log 2>&1; echo "rc=$?"; grep -E "^step|DIST_" gpurun_out/dist$W.log | cut -c1-260 | tail -8
run_bench() { # name nprocs args...
  local name=$1 np=$2; shift 2
  if [ "$np" = 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/scale_$name.json 2> gpurun_out/scale_$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $np "$@" > gpurun_out/scale_$name.json 2> gpurun_out/scale_$name.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$name.json").read().strip().splitlines()[-1])
    print("%-22s n_gpus %d total %d value %.4g ms/step %.3f e2e %s stage %s" % ("$name", d["n_gpus"], d["config"]["particles_total"], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/scale_$name.err").read()[-1500:])
PY
}
for np in 1 2 4 8; do [ $np -le $NG ] && run_bench cfg2_n$np $np --steps 50 --warmup 10 --e2e-steps 5 --no-cpu-baseline; done
[ $NG -ge 8 ] && run_bench cfg4_river16m_n8 8 --config config4_river_16m --particles 2097152 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline
[ $NG -ge 8 ] && run_bench sweep32m_n8 8 --config sweep_4m --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline
echo done
