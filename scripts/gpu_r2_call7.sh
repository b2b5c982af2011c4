#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tail_and_reduce.py tests/test_e2e_parity.py -m gpu -q 2>&1 | tail -5
for v in "1 1" "0 1" "1 0" "0 0" "1 1"; do
set -- $v
CAST_FORK_REDUCE=$1 CAST_FUSE_EMBED_BWD=$2 timeout 600 python bench.py --steps 100 --warmup 5 --no_cpu_baseline --no_eval > gpurun_out/ab_$1$2.json 2> gpurun_out/ab_$1$2.err
echo "fork=$1 embedfx=$2: $(python scripts/show_bench.py gpurun_out/ab_$1$2.json 2>/dev/null | head -1 | cut -c1-60)"
done
