#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CAST_ATTN_KG=4 timeout 900 python -m pytest tests/test_attention_mma.py tests/test_e2e_parity.py tests/test_baseline_shapes.py -m gpu -q -x 2>&1 | tail -3
for kg in 4 2; do
CAST_ATTN_KG=$kg timeout 600 python bench.py --steps 100 --warmup 5 --no_cpu_baseline --no_eval > gpurun_out/r2_bench_c2_kg$kg.json 2> gpurun_out/r2_bench_c2_kg$kg.err
echo "KG=$kg"; python scripts/show_bench.py gpurun_out/r2_bench_c2_kg$kg.json 2>/dev/null | grep -v roofline | head -7
done
