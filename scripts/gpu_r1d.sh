mkdir -p gpurun_out
export OMP_NUM_THREADS=4
timeout 600 python -m pytest tests/test_gpu_homog.py tests/test_gpu_noise.py -m gpu -q -rf -n 4 > gpurun_out/pytest_new3.log 2>&1; echo "new rc=$?"
tail -5 gpurun_out/pytest_new3.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_q2.json 2> gpurun_out/bench_q2.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_q2.json 2>/dev/null | head -3
