#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_6.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_6.txt
tail -12 gpurun_out/r2_pytest_gpu_6.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no_cpu_baseline --no_eval > gpurun_out/r2_bench_c2_d.json 2> gpurun_out/r2_bench_c2_d.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/r2_bench_c2_d.json 2>/dev/null | grep -v roofline
for c in c3 c5; do
timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no_cpu_baseline --no_eval > gpurun_out/r2_bench_${c}_d.json 2> gpurun_out/r2_bench_${c}_d.err; echo "bench $c rc=$?"
python scripts/show_bench.py gpurun_out/r2_bench_${c}_d.json 2>/dev/null | head -2
done
