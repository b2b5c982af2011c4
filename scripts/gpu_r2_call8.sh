#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rowk.py tests/test_e2e_parity.py tests/test_baseline_shapes.py -m gpu -q 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 100 --warmup 5 --no_cpu_baseline --no_eval > gpurun_out/r2_bench_c2_e$i.json 2> gpurun_out/r2_bench_c2_e$i.err
python scripts/show_bench.py gpurun_out/r2_bench_c2_e$i.json 2>/dev/null | grep -v roofline | head -14
done
