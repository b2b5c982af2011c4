#!/bin/bash
# A/B of programmatic dependent launch (CAST_PDL) on the C2 and C1 steps, after the engine-level GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_parity.py tests/test_baseline_shapes.py tests/test_attention_mma.py -m gpu -q -x 2>&1 | tail -1
for p in 1 0 1; do
CAST_PDL=$p timeout 600 python bench.py --steps 100 --warmup 5 --no_cpu_baseline --no_eval > gpurun_out/r2_bench_c2_pdl$p.json 2> gpurun_out/r2_bench_c2_pdl$p.err
echo "PDL=$p $(python scripts/show_bench.py gpurun_out/r2_bench_c2_pdl$p.json 2>/dev/null | head -1 | cut -c1-120)"
done
CAST_PDL=1 timeout 600 python bench.py --config c1 --steps 100 --warmup 5 --no_cpu_baseline --no_eval 2>/dev/null | python scripts/show_bench.py /dev/stdin 2>/dev/null | head -1 | cut -c1-120
