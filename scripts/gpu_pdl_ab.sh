cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for p in 1 0 1 0; do
CAST_PDL=$p timeout 600 python bench.py --steps 100 --warmup 5 --no_cpu_baseline --no_eval > gpurun_out/r2_bench_c2_pdl$p.json 2> gpurun_out/r2_bench_c2_pdl$p.err
echo "PDL=$p $(python scripts/show_bench.py gpurun_out/r2_bench_c2_pdl$p.json 2>/dev/null | head -1 | cut -c1-120)"
done
for p in 1 0; do
CAST_PDL=$p timeout 600 python bench.py --config c1 --steps 100 --warmup 5 --no_cpu_baseline --no_eval 2>/dev/null | python scripts/show_bench.py /dev/stdin 2>/dev/null | head -1 | cut -c1-120
done
