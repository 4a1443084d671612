mkdir -p gpurun_out
export OMP_NUM_THREADS=4
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_homog.py tests/test_gpu_fullsize.py -m gpu -q -rf -n 4 -x > gpurun_out/pytest_heads2.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/pytest_heads2.log
timeout 300 python bench.py --workload lsun64_b256_t16 --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_lsun_v3.json 2> gpurun_out/bench_lsun_v3.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_lsun_v3.json 2>/dev/null | head -6
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_q3.json 2> gpurun_out/bench_q3.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_q3.json 2>/dev/null | head -6
