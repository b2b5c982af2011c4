#!/bin/bash
# GPU-box profiling pass (B200_PROFILING.md recipe): plain run first, then the ncu launch list, then one
# `--set full` capture of the heavy kernels.  Outputs under gpurun_out/ (copied into profiles/ by hand).
TAG=${1:-r01}
CMD="python bench.py --steps 3 --warmup 3 --no_cpu_baseline --no_profile"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 500 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-attn_|qkv_bwd|ffn_bwd}" -s ${KSKIP:-24} -c ${KCOUNT:-8} \
    -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/ncu_full_$TAG.log
