#!/bin/bash
# N-GPU scaling evidence on one box: C2 (default bench, the driver's launch line) and C5 (row-sharded table update)
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_c2_n$N.log 2>&1
echo "c2 rc=$?"; grep '^{' gpurun_out/bench_c2_n$N.log | tail -1 | cut -c1-500
bash scripts/gpu_c5_scale.sh $N
