"""bf16 smoke configuration: per-tensor gradient error against the oracle (which tensors are off?)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import seqvae_b200 as S
from oracle import seqvae_oracle as O
over = dict(filter_sizes=[3, 16, 32, 32, 48, 48], mc_steps=2)
ds = S.SyntheticDataset("celebA", 4, data_dims=[32, 32, 3], data_range=[-1.0, 1.0])
B = 4
model = S.SequentialVAE(ds, B, "c_inhomog", operand_dtype="bf16", restore=False, **over)
hp = O.hyperparams("c_inhomog", [32, 32, 3], (-1.0, 1.0), **over)
P = O.init_params(hp, 0)
model.set_params({k: v.numpy() for k, v in P.items()})
g = torch.Generator().manual_seed(0)
x = (torch.rand(B, 32, 32, 3, generator=g, dtype=torch.float64) * 2 - 1)
eps = torch.randn(2, B, hp["latent_dim"], generator=g, dtype=torch.float64)
if os.environ.get("ORACLE_EMU", "0") == "1":
    from seqvae_b200 import _cabi
    L = _cabi.lib()
    def pred(direction):
        return lambda kind, h, w, cin, cout, stride: bool(
            L.svae_op_tc_supported({"conv": 0, "deconv": 1, "fc": 2}[kind], h, w, cin, cout, stride, direction))
    O.OPERAND_EMULATION = dict(fwd=pred(0), dgrad=pred(1), wgrad=pred(2))
fw, grads = O.loss_and_grads(hp, P, x, x, eps, 0.5)
out = model.forward(x.numpy(), None, eps.numpy(), 0.5)
model.backward()
G = model.gradients(live_only=True)
print("RESULT max|x-x_ref| = %.3e  loss %.6f vs %.6f" % (float(np.abs(out["x"] - torch.stack(fw["x"]).numpy()).max()), out["loss"], float(fw["loss"])))
errs = []
for k, gv in G.items():
    ref = grads[k].numpy()
    errs.append((float(np.linalg.norm(gv - ref) / (np.linalg.norm(ref) + 1e-12)), k, float(np.linalg.norm(ref))))
print("forward: mu err %.2e  sigma err %.2e  x per step %s" % (
    float(np.abs(out["mu"] - torch.stack(fw["mu"]).numpy()).max()), float(np.abs(out["sigma"] - torch.stack(fw["sigma"]).numpy()).max()),
    [float(np.abs(out["x"][t] - fw["x"][t].numpy()).max()) for t in range(2)]))
print("in parameter order:")
for e, k, n in errs:
    if "biases" in k and "fully_connected" not in k: continue
    print("  %.2e %s" % (e, k))
errs.sort(reverse=True)
for e, k, n in errs[:12]:
    print("  %.3e  |ref|=%.3e  %s" % (e, n, k))
print("tensors with err > 0.15: %d of %d" % (sum(1 for e in errs if e[0] > 0.15), len(errs)))
