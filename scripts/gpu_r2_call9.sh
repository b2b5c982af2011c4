#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 100 --warmup 5 --no_cpu_baseline --no_eval > gpurun_out/r2_bench_c2_f$i.json 2> gpurun_out/r2_bench_c2_f$i.err
python scripts/show_bench.py gpurun_out/r2_bench_c2_f$i.json 2>/dev/null | grep -v roofline | head -14
done
