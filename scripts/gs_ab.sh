python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; tail -4 gpurun_out/pytest_q.log
for n in 2 4 8; do
SVAE_GRAD_SETS=$n python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_gs$n.json 2> gpurun_out/bench_gs$n.err || tail -c 800 gpurun_out/bench_gs$n.err
echo "GRAD_SETS=$n"; python scripts/show_bench.py gpurun_out/bench_gs$n.json | head -1
done
nvidia-smi --query-gpu=memory.used --format=csv
