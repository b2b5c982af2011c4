#!/bin/bash
# 2-GPU checks: replicated C2 and row-sharded C4 / C5 with on-box parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r2_mgpu_${name}_n$N.json 2> gpurun_out/r2_mgpu_${name}_n$N.err; echo "$name rc=$?"; tail -3 gpurun_out/r2_mgpu_${name}_n$N.err | cut -c1-400; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_mgpu_${name}_n$N.json") if l.startswith("{")][-1])
    print({k:d.get(k) for k in ("value","ms_per_step","e2e","dp_parity","eval","notes")})
except Exception as e: print("no json", e)
PY
}
[ -n "$SKIP_C2" ] || run c2 --steps 20 --warmup 5 --no_profile --no_cpu_baseline
run c4shard --config c4 --shard_item_table --steps 10 --warmup 3 --no_profile --no_cpu_baseline
run c5 --config c5 --steps 5 --warmup 3 --no_profile --no_cpu_baseline
