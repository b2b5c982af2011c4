#!/bin/bash
# GPU-box check after a row-kernel change: parity tests, then short benches with the FFMA (0) and tensor-core (1) row kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for be in ${BACKENDS:-0 1}; do
  export CAST_FUSED_BACKEND=$be
  timeout 300 python bench.py --steps 20 --warmup 3 --no_cpu_baseline --no_eval > gpurun_out/bench_fused_$be.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_fused_$be.log
  python scripts/show_bench.py gpurun_out/bench_fused_$be.log 2>&1 | grep -v "^roofline\|^cpu" | head -10
done
