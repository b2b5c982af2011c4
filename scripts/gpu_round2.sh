#!/bin/bash
# Round profile pass: other configs sanity (short), launch list + full-set captures of the top kernels at C2
TAG=${1:-r02e}
mkdir -p gpurun_out
CONFIGS="c1 c3 c4 c5" bash scripts/gpu_configs.sh 2>&1 | grep -v "^$" | grep "==\|rc=\|value"
KREGEX="attn_|qkv_bwd|ffn_bwd|ln_qkv|ln_ffn" KSKIP=14 KCOUNT=14 bash scripts/gpu_profile.sh $TAG
