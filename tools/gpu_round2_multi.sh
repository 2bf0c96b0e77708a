#!/usr/bin/env bash
# Round 2, multi-GPU call (gpurun --gpus 2, 4 or 8): the slab decomposition in both organisations, then the
# scaling series. The sub-cell order (--sub) additionally asserts the merged global array ORDER and BITWISE
# equality with a single-GPU run.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_round2_multi.sh'
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
for sub in "" "--sub"; do
  tag=${sub:+_sub}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29631 \
      tests/dist_worker.py $((NG*60000)) 6 $sub > gpurun_out/r02_dist${NG}${tag}.log 2>&1
  echo "dist worker world=$NG $sub rc=$?"; grep -E "^step|DIST_" gpurun_out/r02_dist${NG}${tag}.log | cut -c1-240 | tail -9
done
for org in default auto; do
  for np in 1 2 4 8; do
    [ $np -le $NG ] || continue
    out=gpurun_out/r02_scale_${org}_n$np
    if [ $np = 1 ]; then timeout 900 python bench.py --gpus 1 --organisation $org --steps 50 --warmup 10 --e2e-steps 5 --no-cpu-baseline > $out.json 2> $out.err
    else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29632 \
        bench.py --gpus $np --organisation $org --steps 50 --warmup 10 --e2e-steps 5 --no-cpu-baseline > $out.json 2> $out.err; fi
    python - <<PY
import json
try:
    d = json.loads(open("$out.json").read().strip().splitlines()[-1])
    print("%-8s n_gpus %d value %.4g ms/step %.3f options %s crosscheck %s stage %s" % ("$org", d["n_gpus"], d["value"], d["ms_per_step"],
          d["config"]["options"], d["config"]["organisation"].get("multi_gpu_crosscheck"), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()}))
except Exception as e:
    print("$org n=$np failed:", e); print(open("$out.err").read()[-1200:])
PY
  done
done
