mkdir -p gpurun_out
export OMP_NUM_THREADS=4
timeout 600 python -m pytest tests/test_gpu_homog.py -m gpu -q -rf -n 4 > gpurun_out/pytest_new4.log 2>&1; echo "new rc=$?"
tail -5 gpurun_out/pytest_new4.log
