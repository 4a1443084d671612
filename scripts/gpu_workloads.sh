# bench lines of the other BASELINE configurations and of the homogeneous chains (profiles/r1f_bench_<workload>.json)
mkdir -p gpurun_out
for w in mnist32_b100 cifar32_b100 lsun64_b256_t16 celeba64_b100_homog celeba64_b100_homog_t25; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
python scripts/show_bench.py gpurun_out/bench_$w.json 2>/dev/null | head -1
done
