import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import seqvae_oracle as O
from gpu_util import dev, op_handle, ptr, rel_err
m, L, h = op_handle()
for (B,H,Ci,Co,stride) in [(2, 16, 3, 8, 2), (3, 8, 8, 8, 1), (2, 2, 16, 24, 2), (2, 32, 32, 32, 1), (2, 8, 128, 128, 1)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H, H, Ci, generator=g, dtype=torch.float64)
    w = torch.randn(4, 4, Ci, Co, generator=g, dtype=torch.float64) * 0.1
    ref = O.conv2d_same(x, w, stride)
    ref32 = O.conv2d_same(x.float(), w.float(), stride)
    y = torch.empty(B, H // stride, H // stride, Co, device="cuda")
    stats = torch.zeros(2 * Co, dtype=torch.float64, device="cuda")
    rc = L.svae_op_conv2d(h, ptr(dev(x)), ptr(dev(w)), ptr(y), ptr(stats), B, H, H, Ci, Co, stride, 0)
    m.sync()
    yy = y.double().cpu()
    print("conv", (B,H,Ci,Co,stride), "rc", rc, "err vs f64 %.3e" % rel_err(y.cpu().numpy(), ref.numpy()), "cpu32 vs f64 %.3e" % rel_err(ref32.numpy(), ref.numpy()),
          "stats sum err %.3e" % float((stats[:Co].cpu() - yy.sum(dim=(0,1,2))).abs().max()), "sq err %.3e" % float((stats[Co:].cpu() - (yy**2).sum(dim=(0,1,2))).abs().max()))
    xg = x.clone().requires_grad_(True); wg = w.clone().requires_grad_(True)
    dy = torch.randn(B, H // stride, H // stride, Co, generator=g, dtype=torch.float64)
    O.conv2d_same(xg, wg, stride).backward(dy)
    dx = torch.empty(B, H, H, Ci, device="cuda"); dw = torch.empty(4, 4, Ci, Co, device="cuda")
    rc = L.svae_op_conv2d_backward(h, ptr(dev(x)), ptr(dev(w)), ptr(dev(dy)), ptr(dx), ptr(dw), B, H, H, Ci, Co, stride, 0)
    m.sync()
    print("   bwd rc", rc, "dx err %.3e" % rel_err(dx.cpu().numpy(), xg.grad.numpy()), "dw err %.3e" % rel_err(dw.cpu().numpy(), wg.grad.numpy()))
