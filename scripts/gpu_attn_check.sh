#!/bin/bash
# GPU-box check after an attention change: parity tests first (bounded), then short benches per chunk width
mkdir -p gpurun_out
for ch in ${CHUNKS:-32 64}; do
  export CAST_ATTN_CHUNK=$ch
  timeout 600 python -m pytest tests/test_attention_mma.py tests/test_e2e_parity.py -m gpu -x -q > gpurun_out/pytest_attn_$ch.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_attn_$ch.log; tail -4 gpurun_out/pytest_attn_$ch.log
  timeout 300 python bench.py --steps 20 --warmup 3 --no_cpu_baseline --no_eval > gpurun_out/bench_attn_$ch.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_attn_$ch.log
  python scripts/show_bench.py gpurun_out/bench_attn_$ch.log 2>&1 | grep -v "^roofline\|^cpu" | head -9
done
