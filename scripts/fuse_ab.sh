python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; tail -6 gpurun_out/pytest_q.log
for p in 0 1; do
SVAE_FUSE=$p python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_fuse$p.json 2> gpurun_out/bench_fuse$p.err || tail -c 800 gpurun_out/bench_fuse$p.err
echo "FUSE=$p"; python scripts/show_bench.py gpurun_out/bench_fuse$p.json > gpurun_out/show_fuse$p.txt; head -1 gpurun_out/show_fuse$p.txt
done
cat gpurun_out/show_fuse1.txt
