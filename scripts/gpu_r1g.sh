mkdir -p gpurun_out
export OMP_NUM_THREADS=4
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_homog.py tests/test_gpu_fullsize.py -m gpu -q -rf -n 4 -k "lsun or homog_v1 or homog" > gpurun_out/pytest_heads.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/pytest_heads.log
timeout 300 python bench.py --workload lsun64_b256_t16 --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_lsun_v2.json 2> gpurun_out/bench_lsun_v2.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_lsun_v2.json 2>/dev/null | head -6
