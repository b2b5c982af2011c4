#!/bin/bash
# C5 (1M-item catalog, hidden 256, maxlen 200, per-GPU batch 128; item table row-sharded over the ranks) at the GPU
# counts given as arguments:  bash scripts/gpu_c5_scale.sh 1 2   /   bash scripts/gpu_c5_scale.sh 4 8
# Each run prints bench.py's JSON line into gpurun_out/scale_c5_n<N>.json (dp_parity, step_split, eval included).
# CFG=c2 (or any other bench config) runs that workload instead: gpurun_out/scale_<CFG>_n<N>.json.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CFG=${CFG:-c5}
STEPS=${STEPS:-10}
for N in "$@"; do
  if [ "$N" = 1 ]; then
    timeout ${TMO:-600} python bench.py --gpus 1 --config $CFG --steps $STEPS --warmup 3 --no_profile --no_cpu_baseline \
      > gpurun_out/scale_${CFG}_n$N.json 2> gpurun_out/scale_${CFG}_n$N.err
  else
    timeout ${TMO:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29517 bench.py --gpus $N --config $CFG --steps $STEPS --warmup 3 --no_profile --no_cpu_baseline \
      > gpurun_out/scale_${CFG}_n$N.json 2> gpurun_out/scale_${CFG}_n$N.err
  fi
  echo "N=$N rc=$?"
  grep '^{' gpurun_out/scale_${CFG}_n$N.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d.get(k) for k in ('value','ms_per_step','step_split','dp_parity','hbm_per_gpu')}); print(d.get('eval'))" 2>/dev/null || tail -5 gpurun_out/scale_${CFG}_n$N.err | cut -c1-300
done
