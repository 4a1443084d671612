"""Per-CTA phase timeline of the tcgen05 conv kernel (globaltimer marks written by the kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, ctypes as C
from gpu_util import op_handle, ptr
m, L, h = op_handle()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
def run(kind, H, Ci, Co, stride):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, H, Ci, device="cuda", generator=g)
    w = torch.randn(4, 4, *( (Ci, Co) if kind == "conv" else (Co, Ci)), device="cuda", generator=g) * 0.05
    Ho = H // stride if kind == "conv" else H * stride
    y = torch.empty(B, Ho, Ho, Co, device="cuda")
    stats = torch.zeros(2 * Co, dtype=torch.float64, device="cuda")
    dbg = torch.zeros(8192 * 16 + 128, dtype=torch.int64, device="cuda")
    fn = L.svae_op_conv2d if kind == "conv" else L.svae_op_conv2d_transpose
    torch.cuda.synchronize()
    for it in range(3):
        L.svae_debug_set_buffer(ptr(dbg) if it == 2 else None)
        assert fn(h, ptr(x), ptr(w), ptr(y), ptr(stats), B, H, H, Ci, Co, stride, 1) == 0
        m.sync()
    L.svae_debug_set_buffer(None)
    d = dbg.cpu().numpy()[:8192 * 16].reshape(-1, 16)
    d = d[d[:, 0] > 0]
    t0 = d[:, 0].min()
    names = ["setup", "produce", "mma(wait acc)", "epilogue", "teardown"]
    print("== %s H=%d %d->%d s%d: %d CTAs, kernel span %.1f us" % (kind, H, Ci, Co, stride, len(d), (d[:, 5].max() - t0) / 1e3))
    for k, nme in enumerate(names):
        seg = (d[:, k + 1] - d[:, k]) / 1e3
        print("   %-14s mean %.2f us  p50 %.2f  p95 %.2f" % (nme, seg.mean(), np.median(seg), np.percentile(seg, 95)))
    print("   mma thread: a_ready seen %.2f us after producer warp0 done; issue loop %.2f us; acc_done seen %.2f us after last issue" % (
        ((d[:, 6] - d[:, 2]) / 1e3).mean(), ((d[:, 7] - d[:, 6]) / 1e3).mean(), ((d[:, 3] - d[:, 7]) / 1e3).mean()))
    print("   producer warps done at (us after setup): " + " ".join("%.2f" % ((d[:, 8 + w] - d[:, 1]) / 1e3).mean() for w in range(4)))
    print("   (us after setup) mma warp enters role %.2f | warp0 staging done (pre-fence) %.2f | warp0 arrived %.2f | mma sees a_ready %.2f" % (
        ((d[:, 12] - d[:, 1]) / 1e3).mean(), ((d[:, 14] - d[:, 1]) / 1e3).mean(), ((d[:, 2] - d[:, 1]) / 1e3).mean(), ((d[:, 6] - d[:, 1]) / 1e3).mean()))
    life = (d[:, 5] - d[:, 0]) / 1e3
    print("   CTA lifetime   mean %.2f us ; start times: p50 %.1f p95 %.1f us" % (life.mean(), np.median(d[:, 0] - t0) / 1e3, np.percentile(d[:, 0] - t0, 95) / 1e3))
run("conv", 32, 32, 32, 1)
run("deconv", 32, 64, 32, 1)
run("conv", 8, 128, 128, 1)
run("deconv", 16, 64, 32, 2)
