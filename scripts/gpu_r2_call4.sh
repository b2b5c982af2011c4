#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowk.py tests/test_baseline_shapes.py -m gpu -x -q 2>&1 | tail -5
python scripts/trace_rowk.py 2>&1 | tail -14
timeout 600 python bench.py --steps 20 --warmup 5 --no_cpu_baseline --no_eval > gpurun_out/r2_bench_c2_b.json 2> gpurun_out/r2_bench_c2_b.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/r2_bench_c2_b.json 2>/dev/null | grep -v roofline
