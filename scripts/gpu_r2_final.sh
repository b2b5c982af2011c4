#!/bin/bash
# final evidence pass of the round (1 GPU): GPU test suite, smoke, the default bench line and the reference arm,
# then scripts/gpu_r2_profiles.sh (other configs, ncu launch list, --set full summaries)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r04}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/${TAG}_bench_c2.json 2>&1 | head -18
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/${TAG}_bench_ref.json | cut -c1-400
bash scripts/gpu_r2_profiles.sh $TAG
