#!/bin/bash
# ncu --set full on kernels matching KREGEX of an arbitrary python command: TAG KREGEX SKIP COUNT -- cmd...
TAG=$1; KREGEX=$2; SKIP=$3; COUNT=$4; shift 5
mkdir -p gpurun_out
"$@" > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s $SKIP -c $COUNT \
    -o gpurun_out/prof_$TAG -f "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full_$TAG.log
