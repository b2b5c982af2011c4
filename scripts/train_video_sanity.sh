#!/bin/bash
# End-to-end sanity run on the one real dataset present in the reference tree (data/Video.txt = Amazon "Games" of the
# SASRec paper: 31 013 users, 23 715 items, 287 107 actions).  The file has the upstream 2-column `user item` layout; the
# reference's util.get_users wants 4 columns, so it is rewritten with rating 1 and per-user increasing timestamps
# (SASRec ignores both).  The converted copy lives under data/_scratch/ (git-ignored, travels with gpurun).
#   build container:  bash scripts/train_video_sanity.sh convert
#   GPU box:          bash scripts/train_video_sanity.sh train
set -e
if [ "$1" = "convert" ]; then
  mkdir -p data/_scratch
  python - <<'PY'
from collections import defaultdict
n = defaultdict(int)
with open('/root/reference/data/Video.txt') as f, open('data/_scratch/Video4.txt', 'w') as g:
    for line in f:
        u, i = line.split()
        n[u] += 1
        g.write(f"{u} {i} 1 {1000000000 + 86400 * n[u]}\n")
print(len(n), "users")
PY
else
  python main.py --dataset data/_scratch/Video4.txt --train_dir video_sanity --model sasrec --maxlen 50 \
    --dropout_rate 0.5 --num_epochs 201 --model_path gpurun_out/saved_models > gpurun_out/video_train.log 2>&1
  tail -3 gpurun_out/video_train.log
  cat gpurun_out/saved_models/*/video_sanity*/log.txt
fi
