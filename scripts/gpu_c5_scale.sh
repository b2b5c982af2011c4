#!/bin/bash
# C5 (1M-item catalog, row-sharded table update) at N GPUs: bash scripts/gpu_c5_scale.sh N
N=${1:-2}
timeout ${TMO:-250} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --config c5 --steps 10 --warmup 3 --no_eval --no_profile > gpurun_out/bench_c5_n$N.log 2>&1
echo rc=$?
grep '^{' gpurun_out/bench_c5_n$N.log | tail -1 | cut -c1-700
grep -v '^{' gpurun_out/bench_c5_n$N.log | tail -5 | cut -c1-300
