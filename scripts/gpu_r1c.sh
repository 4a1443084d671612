mkdir -p gpurun_out
export OMP_NUM_THREADS=4
timeout 600 python -m pytest tests/test_gpu_homog.py tests/test_gpu_noise.py -m gpu -q -rf -n 4 > gpurun_out/pytest_new2.log 2>&1; echo "new rc=$?"
tail -5 gpurun_out/pytest_new2.log
for w in celeba64_b100_homog celeba64_b100_homog_t25; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
python scripts/show_bench.py gpurun_out/bench_$w.json | head -8
done
