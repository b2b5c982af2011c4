#!/bin/bash
# ncu --set full on selected kernels of one bench configuration: TAG KREGEX CONFIG [SKIP] [COUNT]
TAG=$1; KREGEX=$2; CFGN=${3:-c2}; SKIP=${4:-20}; COUNT=${5:-6}
CMD="python bench.py --config $CFGN --steps 3 --warmup 3 --no_cpu_baseline --no_profile --no_eval"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s $SKIP -c $COUNT \
    -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full_$TAG.log
