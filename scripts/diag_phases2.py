"""Per-CTA phase timeline of the TMA-fed persistent conv kernel (tc2): globaltimer marks written by the kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gpu_util import op_handle, ptr
m, L, h = op_handle()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
def run(kind, H, Ci, Co, stride):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, H, Ci, device="cuda", generator=g)
    w = torch.randn(4, 4, *((Ci, Co) if kind == "conv" else (Co, Ci)), device="cuda", generator=g) * 0.05
    Ho = H // stride if kind == "conv" else H * stride
    y = torch.empty(B, Ho, Ho, Co, device="cuda")
    stats = torch.zeros(2 * Co, dtype=torch.float64, device="cuda")
    dbg = torch.zeros(8192 * 16 + 128, dtype=torch.int64, device="cuda")
    fn = L.svae_op_conv2d if kind == "conv" else L.svae_op_conv2d_transpose
    torch.cuda.synchronize()
    NW = int(os.environ.get("DIAG_WARM", "3"))     # warm-up launches (GPU clocks ramp up under load)
    for it in range(NW):
        L.svae_debug_set_buffer(ptr(dbg) if it == NW - 1 else None)
        assert fn(h, ptr(x), ptr(w), ptr(y), ptr(stats), B, H, H, Ci, Co, stride, 1) == 0
        m.sync()
    L.svae_debug_set_buffer(None)
    d = dbg.cpu().numpy()[:8192 * 16].reshape(-1, 16)
    d = d[d[:, 0] > 0]
    t0 = d[:, 0].min()
    print("== %s H=%d %d->%d s%d: %d CTAs, tiles/CTA %.1f, kernel span %.1f us; CTA start p50 %.1f p95 %.1f us" % (
        kind, H, Ci, Co, stride, len(d), d[:, 15].mean(), (d[:, 10].max() - t0) / 1e3,
        np.median(d[:, 0] - t0) / 1e3, np.percentile(d[:, 0] - t0, 95) / 1e3))
    def seg(a, b): return ((d[:, b] - d[:, a]) / 1e3).mean()
    print("   setup %.2f | first halo issued +%.2f | weights resident seen +%.2f | first halo landed +%.2f (all after setup)" % (
        seg(0, 1), seg(1, 2), seg(1, 3), seg(1, 4)))
    print("   tile0: mma issue+commit %.2f | epilogue sees acc +%.2f after commit | epilogue %.2f" % (seg(4, 5), seg(5, 6), seg(6, 7)))
    print("   last mma commit at %.2f, last epilogue done at %.2f, exit at %.2f (us after setup)" % (seg(1, 8), seg(1, 9), seg(1, 10)))
cases = [("conv", 32, 32, 32, 1), ("deconv", 32, 64, 32, 1), ("conv", 8, 128, 128, 1), ("deconv", 16, 64, 32, 2), ("conv", 16, 64, 64, 1)]
if len(sys.argv) > 2 and sys.argv[2] == "small":
    cases = [("conv", 8, 128, 128, 2), ("deconv", 4, 128, 128, 2), ("deconv", 4, 384, 128, 2), ("deconv", 8, 256, 128, 1), ("conv", 8, 128, 128, 1)]
for s in cases:
    run(*s)
