"""Process-isolated check of the tcgen05 kernels against the oracle (one subprocess per shape, bounded time)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    # kind, B, H, Ci, Co, stride
    ("conv", 2, 8, 16, 16, 1), ("conv", 2, 16, 32, 32, 1), ("conv", 3, 16, 64, 64, 1), ("conv", 2, 8, 128, 128, 1),
    ("conv", 2, 16, 32, 64, 2), ("conv", 4, 8, 128, 128, 2), ("conv", 100, 32, 32, 32, 1),
    ("deconv", 2, 8, 32, 16, 1), ("deconv", 3, 8, 256, 128, 1), ("deconv", 2, 16, 64, 32, 2), ("deconv", 2, 4, 384, 128, 2),
    ("deconv", 100, 16, 64, 32, 2), ("deconv", 100, 32, 64, 32, 1),
]
if len(sys.argv) > 1:
    CASES = CASES[:int(sys.argv[1])]
CHILD = r'''
import sys, os
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch
from oracle import seqvae_oracle as O
from gpu_util import dev, op_handle, ptr, rel_err
kind, B, H, Ci, Co, stride = %(case)r
m, L, h = op_handle()
g = torch.Generator().manual_seed(1)
x = torch.randn(B, H, H, Ci, generator=g, dtype=torch.float64)
xb = x.float().bfloat16().double()
if kind == "conv":
    w = torch.randn(4, 4, Ci, Co, generator=g, dtype=torch.float64) * 0.1
    ref = O.conv2d_same(x, w, stride); refb = O.conv2d_same(xb, w.float().bfloat16().double(), stride)
    y = torch.empty(B, H // stride, H // stride, Co, device="cuda"); fn = L.svae_op_conv2d
else:
    w = torch.randn(4, 4, Co, Ci, generator=g, dtype=torch.float64) * 0.1
    ref = O.conv2d_transpose_same(x, w, stride); refb = O.conv2d_transpose_same(xb, w.float().bfloat16().double(), stride)
    y = torch.empty(B, H * stride, H * stride, Co, device="cuda"); fn = L.svae_op_conv2d_transpose
y.fill_(float("nan"))
stats = torch.zeros(2 * Co, dtype=torch.float64, device="cuda")
ux, uw = dev(x), dev(w)
torch.cuda.synchronize()
rc = fn(h, ptr(ux), ptr(uw), ptr(y), ptr(stats), B, H, H, Ci, Co, stride, 1)
if rc != 0:
    print("rc", rc, L.svae_last_error(h).decode()); sys.exit(0)
m.sync()
yy = y.double().cpu()
nan = int(torch.isnan(yy).sum())
print("err vs bf16-rounded oracle %%.3e | vs fp64 %%.3e | nan %%d | stats err %%.2e %%.2e" %% (
    rel_err(torch.nan_to_num(yy).numpy(), refb.numpy()), rel_err(torch.nan_to_num(yy).numpy(), ref.numpy()), nan,
    float((stats[:Co].cpu() - torch.nan_to_num(yy).sum(dim=(0,1,2))).abs().max()),
    float((stats[Co:].cpu() - (torch.nan_to_num(yy)**2).sum(dim=(0,1,2))).abs().max())))
'''
for case in CASES:
    code = CHILD % dict(root=ROOT, case=case)
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
        out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
        if r.returncode != 0:
            out += " | EXIT %d: %s" % (r.returncode, (r.stderr.strip().splitlines() or [""])[-1][:200])
    except subprocess.TimeoutExpired:
        out = "TIMEOUT"
    print(case, "->", out, flush=True)
