#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowk.py -m gpu -x -q > gpurun_out/r2_pytest_rowk.txt 2>&1; echo "rowk rc=$?" >> gpurun_out/r2_pytest_rowk.txt
tail -5 gpurun_out/r2_pytest_rowk.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_3.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_3.txt
tail -8 gpurun_out/r2_pytest_gpu_3.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/r2_bench_c2_a.json 2> gpurun_out/r2_bench_c2_a.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/r2_bench_c2_a.json 2>/dev/null || tail -c 3000 gpurun_out/r2_bench_c2_a.json
