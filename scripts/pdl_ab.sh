python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; tail -4 gpurun_out/pytest_q.log
for p in 0 1; do
SVAE_PDL=$p python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_pdl$p.json 2> gpurun_out/bench_pdl$p.err || tail -c 800 gpurun_out/bench_pdl$p.err
echo "PDL=$p"; python scripts/show_bench.py gpurun_out/bench_pdl$p.json | head -2
done
