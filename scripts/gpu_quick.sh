# quick GPU check (a few minutes of box time): parity suites on 4 workers, short bench, smoke; "ncu" adds the launch list
mkdir -p gpurun_out
export OMP_NUM_THREADS=4
timeout 1500 python -m pytest tests -m gpu -q -rf -n 4 --durations=8 > gpurun_out/pytest_q.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/pytest_q.log
unset OMP_NUM_THREADS
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err || tail -c 800 gpurun_out/bench_q.err
python scripts/show_bench.py gpurun_out/bench_q.json | head -16
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_q.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_q.log
if [ "$1" = "ncu" ]; then
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/plain_q.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1400 --csv --log-file gpurun_out/launches_q.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/ncu_list_q.log 2>&1
fi
