for n in 4 6; do
SVAE_W_STAGES=$n timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_ws$n.json 2> gpurun_out/bench_ws$n.err || tail -c 800 gpurun_out/bench_ws$n.err
echo "W_STAGES=$n"; python scripts/show_bench.py gpurun_out/bench_ws$n.json > gpurun_out/show_ws$n.txt; head -1 gpurun_out/show_ws$n.txt; grep "gather" gpurun_out/show_ws$n.txt
done
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
