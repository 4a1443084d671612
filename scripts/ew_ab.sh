for n in 8 4 2; do
SVAE_EW_CAP=$n timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_ew$n.json 2> gpurun_out/bench_ew$n.err || tail -c 800 gpurun_out/bench_ew$n.err
echo "EW_CAP=$n"; python scripts/show_bench.py gpurun_out/bench_ew$n.json > gpurun_out/show_ew$n.txt; head -1 gpurun_out/show_ew$n.txt; grep "bn_\|out_mix" gpurun_out/show_ew$n.txt
done
