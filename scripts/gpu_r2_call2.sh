#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/debug_c3.py > gpurun_out/r2_debug_c3.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_2.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_2.txt
cat gpurun_out/r2_debug_c3.txt; tail -15 gpurun_out/r2_pytest_gpu_2.txt
