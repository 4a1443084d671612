"""Tiny driver for ncu: a few launches of the main tcgen05 kernels at benchmark shapes (CelebA B=100)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import op_handle, ptr
m, L, h = op_handle()
B = 100
def run(kind, H, Ci, Co, stride, reps=3):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, H, Ci, device="cuda", generator=g)
    if kind == "conv":
        w = torch.randn(4, 4, Ci, Co, device="cuda", generator=g) * 0.05
        Ho = H // stride
    else:
        w = torch.randn(4, 4, Co, Ci, device="cuda", generator=g) * 0.05
        Ho = H * stride
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
run("conv", 32, 32, 32, 1)      # c0b
run("deconv", 32, 64, 32, 1)    # t0b
print("ok")
