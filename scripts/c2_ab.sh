for n in 0 1; do
SVAE_CONV_2CTA=$n timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_c2$n.json 2> gpurun_out/bench_c2$n.err || tail -c 800 gpurun_out/bench_c2$n.err
echo "CONV_2CTA=$n"; python scripts/show_bench.py gpurun_out/bench_c2$n.json > gpurun_out/show_c2$n.txt; head -1 gpurun_out/show_c2$n.txt; grep "gather" gpurun_out/show_c2$n.txt
done
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
DIAG_WARM=20 timeout 100 python scripts/diag_phases2.py 100 2>&1 | sed -n 5,8p
