python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; tail -4 gpurun_out/pytest_q.log
for n in 0 1; do
SVAE_WGRAD_2CTA=$n python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_w2$n.json 2> gpurun_out/bench_w2$n.err || tail -c 800 gpurun_out/bench_w2$n.err
echo "WGRAD_2CTA=$n"; python scripts/show_bench.py gpurun_out/bench_w2$n.json > gpurun_out/show_w2$n.txt; head -1 gpurun_out/show_w2$n.txt; grep "wgrad\|skinny\|pack" gpurun_out/show_w2$n.txt
done
