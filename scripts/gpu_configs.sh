#!/bin/bash
# first look at the other BASELINE configs on one GPU (short runs)
for c in ${CONFIGS:-c1 c3 c4 c5}; do
  echo "== $c"
  timeout 280 python bench.py --config $c --steps 10 --warmup 3 --no_cpu_baseline --no_eval > gpurun_out/bench_$c.log 2>&1; echo rc=$?
  python scripts/show_bench.py gpurun_out/bench_$c.log 2>&1 | grep -v "^roofline\|^cpu" | head -12
  tail -3 gpurun_out/bench_$c.log | grep -v "^{" | cut -c1-300
done
