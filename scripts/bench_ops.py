"""Per-kernel timings of the contraction kernels at the benchmark shapes (CelebA-64, B=100), through the layer-level C ABI
with the library's event profiler on (SVAE_TRACE=1 prints one line per launch).  usage: SVAE_TRACE=1 python scripts/bench_ops.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gpu_util import op_handle, ptr
m, L, h = op_handle()
B = int(os.environ.get("B", "100"))
SHAPES = [("conv", 64, 3, 32, 2), ("conv", 32, 32, 32, 1), ("conv", 32, 32, 64, 2), ("conv", 16, 64, 64, 1),
          ("conv", 16, 64, 128, 2), ("conv", 8, 128, 128, 1), ("conv", 8, 128, 128, 2),
          ("deconv", 4, 384, 128, 2), ("deconv", 8, 256, 128, 1), ("deconv", 8, 128, 64, 2), ("deconv", 16, 128, 64, 1),
          ("deconv", 16, 64, 32, 2), ("deconv", 32, 64, 32, 1), ("deconv", 32, 32, 3, 2)]
def run(kind, H, Ci, Co, stride, reps=3):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, H, Ci, device="cuda", generator=g)
    if kind == "conv":
        w = torch.randn(4, 4, Ci, Co, device="cuda", generator=g) * 0.05; Ho = H // stride
    else:
        w = torch.randn(4, 4, Co, Ci, device="cuda", generator=g) * 0.05; Ho = H * stride
    y = torch.empty(B, Ho, Ho, Co, device="cuda")
    dy = torch.randn(B, Ho, Ho, Co, device="cuda", generator=g)
    dx = torch.empty_like(x); dw = torch.empty_like(w)
    stats = torch.zeros(2 * Co, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(reps):
        if kind == "conv":
            assert L.svae_op_conv2d(h, ptr(x), ptr(w), ptr(y), ptr(stats), B, H, H, Ci, Co, stride, 1) == 0
            assert L.svae_op_conv2d_backward(h, ptr(x), ptr(w), ptr(dy), ptr(dx), ptr(dw), B, H, H, Ci, Co, stride, 1) == 0
        else:
            assert L.svae_op_conv2d_transpose(h, ptr(x), ptr(w), ptr(y), ptr(stats), B, H, H, Ci, Co, stride, 1) == 0
            assert L.svae_op_conv2d_transpose_backward(h, ptr(x), ptr(w), ptr(dy), ptr(dx), ptr(dw), B, H, H, Ci, Co, stride, 1) == 0
    m.sync()
m.profile(True)
for s in SHAPES:
    run(*s)
m.profile_read()
print("ok")
