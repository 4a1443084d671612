# Second-half round-1 GPU check (one B200): the new parity suites (homogeneous chains, device-side denoising corruption,
# full-size configs), then the whole GPU suite, then a short bench line.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
export OMP_NUM_THREADS=4
timeout 600 python -m pytest tests/test_gpu_homog.py tests/test_gpu_noise.py tests/test_gpu_fullsize.py -m gpu -q -rf -n 4 \
  --durations=10 > gpurun_out/pytest_new.log 2>&1; echo "new rc=$?"
tail -40 gpurun_out/pytest_new.log
timeout 600 python -m pytest tests -m gpu -q -rf -n 4 --durations=10 \
  --deselect tests/test_gpu_homog.py --deselect tests/test_gpu_noise.py --deselect tests/test_gpu_fullsize.py \
  > gpurun_out/pytest_old.log 2>&1; echo "old rc=$?"
tail -15 gpurun_out/pytest_old.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_q.json
