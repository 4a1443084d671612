# quick GPU check: parity tests, short bench, per-kernel launch list (serialised) of one graph replay
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; tail -4 gpurun_out/pytest_q.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err || tail -c 800 gpurun_out/bench_q.err
python scripts/show_bench.py gpurun_out/bench_q.json
if [ "$1" = "ncu" ]; then
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/plain_q.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1400 --csv --log-file gpurun_out/launches_q.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/ncu_list_q.log 2>&1
fi
