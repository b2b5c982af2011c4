#!/bin/bash
# experiment: attention tile shapes (CAST_ATT_TUNE bit 0: fwd 32-row tiles, bit 1: dq, bit 2: dkv)
for t in 0 1 2 4 7; do
  echo "== CAST_ATT_TUNE=$t"
  CAST_ATT_TUNE=$t python bench.py --steps 20 --warmup 3 --no_cpu_baseline > gpurun_out/tune_$t.log 2>&1
  python scripts/show_bench.py gpurun_out/tune_$t.log | grep -E "value|attn"
done
