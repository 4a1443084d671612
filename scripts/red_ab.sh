timeout 400 python -m pytest tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -2
for n in 1 2; do
SVAE_RED_CAP=$n timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_red$n.json 2> gpurun_out/bench_red$n.err || tail -c 800 gpurun_out/bench_red$n.err
echo "RED_CAP=$n"; python scripts/show_bench.py gpurun_out/bench_red$n.json > gpurun_out/show_red$n.txt; head -1 gpurun_out/show_red$n.txt; grep "bn_bwd" gpurun_out/show_red$n.txt
done
