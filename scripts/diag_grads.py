"""Diagnostic: per-tensor gradient error of the CUDA path vs the fp64 oracle, next to the error of an fp32 CPU oracle."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import seqvae_oracle as O
from gpu_util import make_pair, make_inputs, rel_err

def run(netname, dims, rng, B, operand, **over):
    model, hp, P = make_pair(netname, dims, rng, B, operand, **over)
    x, eps = make_inputs(hp, B)
    tgt = (x * 0.9).float().double()
    fw, grads = O.loss_and_grads(hp, P, x, tgt, eps, 0.6)
    P32 = {k: v.float() for k, v in P.items()}
    fw32, g32 = O.loss_and_grads(hp, P32, x.float(), tgt.float(), eps.float(), 0.6)
    out = model.forward(x.numpy(), tgt.numpy(), eps.numpy(), 0.6)
    model.backward()
    G = model.gradients(live_only=True)
    rows = []
    for k, gv in G.items():
        ref = grads[k].numpy()
        rows.append((rel_err(gv, ref), rel_err(g32[k].numpy(), ref), k, float(np.linalg.norm(ref))))
    rows.sort(reverse=True)
    print("==", netname, dims, B, operand, "fwd x err", float(np.abs(out["x"] - torch.stack(fw["x"]).numpy()).max()),
          "fp32-oracle x err", float((torch.stack(fw32["x"]).double() - torch.stack(fw["x"])).abs().max()))
    for r in rows[:12]:
        print("  gpu %.2e  cpu32 %.2e  %s |g|=%.2e" % r)
    model.close()

run("m_inhomog", [32, 32, 1], (0.0, 1.0), 6, "fp32", mc_steps=2)
run("m_inhomog", [32, 32, 1], (0.0, 1.0), 32, "fp32", mc_steps=2)
run("c_inhomog", [64, 64, 3], (-1.0, 1.0), 3, "fp32", mc_steps=2)
