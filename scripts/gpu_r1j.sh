mkdir -p gpurun_out
export OMP_NUM_THREADS=4
timeout 600 python -m pytest tests -m gpu -q -rf -n 4 > gpurun_out/pytest_final3.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/pytest_final3.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_final3.json 2>/dev/null | head -1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke3.log
