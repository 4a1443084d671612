set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest17.log
python bench.py > gpurun_out/bench_r1_m.json 2> gpurun_out/bench_r1_m.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/plain_m.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1400 --csv --log-file gpurun_out/launches_r1m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/ncu_list_m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc2_conv_kernel -s 40 -c 3 -o gpurun_out/prof_tc2_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/ncu_full_m.log 2>&1
tail -3 gpurun_out/pytest17.log
