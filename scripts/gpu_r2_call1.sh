#!/bin/bash
# round-2 call 1: hardware probe for MN-major tcgen05 operands + the whole GPU test suite incl. BASELINE-shape parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 scripts/probes/umma_mn_probe > gpurun_out/r2_umma_mn_probe.txt 2>&1; echo "probe rc=$?" >> gpurun_out/r2_umma_mn_probe.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_1.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_1.txt
tail -5 gpurun_out/r2_umma_mn_probe.txt; tail -15 gpurun_out/r2_pytest_gpu_1.txt
