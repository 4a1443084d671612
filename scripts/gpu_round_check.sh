# Round-end evidence run on one B200: parity tests, the default bench line, the ncu launch list of one graph-replayed step and
# one `ncu --set full` capture of the dominant kernel.  Outputs land in gpurun_out/ (copied into profiles/ by hand).
set -x
OMP_NUM_THREADS=4 python -m pytest tests -m gpu -x -q -n 4 > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/ncu_list_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc2_conv_kernel -s 40 -c 12 -o gpurun_out/prof_tc2_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/ncu_full_final.log 2>&1
tail -3 gpurun_out/pytest_final.log
