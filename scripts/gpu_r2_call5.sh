#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_5.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_5.txt
tail -12 gpurun_out/r2_pytest_gpu_5.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no_cpu_baseline --no_eval > gpurun_out/r2_bench_c2_c.json 2> gpurun_out/r2_bench_c2_c.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/r2_bench_c2_c.json 2>/dev/null | grep -v roofline
