#!/bin/bash
# Round profile pass: other configs sanity (short), launch list + full-set captures of the top kernels at C2
mkdir -p gpurun_out
CONFIGS="c1 c3 c4 c5" bash scripts/gpu_configs.sh 2>&1 | grep -v "^$"
KREGEX="attn_|qkv_bwd|ffn_bwd|ln_qkv|ln_ffn" KSKIP=14 KCOUNT=7 bash scripts/gpu_profile.sh r02c
