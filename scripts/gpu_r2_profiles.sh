#!/bin/bash
# round-2 evidence pass (1 GPU): bench lines of the other BASELINE configs, ncu launch list of the C2 step, one
# `--set full` capture of the heavy kernels.  Everything lands in gpurun_out/ (summaries are copied to profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r03}
for spec in "c1" "c3" "c3 --model cast_4" "c3 --model cast_9" "c4" "c5"; do
  name=$(echo $spec | tr -d ' -' | sed 's/model//')
  timeout 400 python bench.py --config $spec --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err
  echo "== $spec rc=$?"; python scripts/show_bench.py gpurun_out/${TAG}_bench_$name.json 2>&1 | grep -v "^roofline\|^cpu" | head -8
done
CMD="python bench.py --steps 3 --warmup 3 --no_cpu_baseline --no_profile --no_eval"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ru_ln|attn_|qkv_bwd|ffn_bwd|segment_|lnf_loss|embed_fwd" -s 40 -c 16 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/${TAG}_ncu_full.log
# only text summaries travel back (the .ncu-rep with sources exceeds the 64 MiB return limit)
python scripts/ncu_summary.py gpurun_out/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.txt 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
rm -f gpurun_out/${TAG}_prof.ncu-rep
ls -la gpurun_out | head -30
