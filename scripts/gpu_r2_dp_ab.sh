#!/bin/bash
# where does a 2-rank C2 step's time go? full exchange vs barrier only vs no exchange, and 1 GPU on the same box
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 200 --warmup 10 --no_eval --no_cpu_baseline > gpurun_out/dpab_n1.json 2> gpurun_out/dpab_n1.err
for m in peer barrier none nccl; do
  CAST_DP_EXCHANGE=$m timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 --no_eval --no_dp_parity > gpurun_out/dpab_$m.json 2> gpurun_out/dpab_$m.err
done
python - <<'PY'
import json
for m in ['n1','peer','barrier','none','nccl']:
    try:
        d=json.loads(open(f'gpurun_out/dpab_{m}.json').read().strip().splitlines()[-1])
        print(m, round(d['value']), round(d['ms_per_step'],4), d['notes'].get('ms_per_step_fastest_rank'), d['e2e']['ms_per_step'], d.get('step_split'))
    except Exception as e:
        print(m,'ERR',e)
PY
