mkdir -p gpurun_out
export OMP_NUM_THREADS=4
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_replay.py tests/test_gpu_fullsize.py -m gpu -q -rf -n 4 -k "production or replay or oracle_parity or probe" > gpurun_out/pytest_p.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/pytest_p.log
