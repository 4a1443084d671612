# A/B: tiled head kernels also for the narrow (CelebA / MNIST) latent groups
mkdir -p gpurun_out
export OMP_NUM_THREADS=4
SVAE_HEADS_TILED=1 timeout 300 python -m pytest tests/test_gpu_chain.py -m gpu -q -n 4 -k "golden or full_depth or trajectory" > gpurun_out/pytest_heads_tiled.log 2>&1; echo "tests rc=$?"
tail -2 gpurun_out/pytest_heads_tiled.log
SVAE_HEADS_TILED=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/bench_heads_tiled.json 2> gpurun_out/bench_heads_tiled.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_heads_tiled.json 2>/dev/null | grep -E "value|skinny"
