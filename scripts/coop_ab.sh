SVAE_COOP_BN=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; tail -4 gpurun_out/pytest_q.log
for n in 0 1; do
SVAE_COOP_BN=$n timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_coop$n.json 2> gpurun_out/bench_coop$n.err || tail -c 800 gpurun_out/bench_coop$n.err
echo "COOP_BN=$n"; python scripts/show_bench.py gpurun_out/bench_coop$n.json > gpurun_out/show_coop$n.txt; head -1 gpurun_out/show_coop$n.txt; grep "bn_" gpurun_out/show_coop$n.txt
done
SVAE_COOP_BN=1 FLOOR_BS=2,100 timeout 200 python scripts/latency_floor.py 2>&1 | tail -2
