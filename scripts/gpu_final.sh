#!/bin/bash
# Round-end pass on one GPU: all gpu tests, smoke, default bench (both arms), then the ncu launch list and one
# --set full capture of the top kernels (after the plain runs exited 0)
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_default.log | head -24
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400
KREGEX="attn_|qkv_bwd|ffn_bwd|ln_qkv|ln_ffn" KSKIP=14 KCOUNT=14 bash scripts/gpu_profile.sh $TAG
