import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import seqvae_oracle as O
from gpu_util import TINY, make_pair, make_inputs, rel_err
over = dict(TINY, mc_steps=3, first_step_loss_coeff=0.5, intermediate_reconstruction=False, regularized_steps=[0, 2],
            latent_mean_clip=0.05, latent_prior_stddev=2.0, min_highway_ratio=0.1, max_highway_ratio=0.8)
model, hp, P = make_pair("c_inhomog", [16, 16, 3], (-1.0, 1.0), 4, "fp32", **over)
hp["regularized_steps"] = [0, 2]
x, eps = make_inputs(hp, 4)
fw, grads = O.loss_and_grads(hp, P, x, x, eps, 0.8)
out = model.forward(x.numpy(), None, eps.numpy(), 0.8)
model.backward()
G = model.gradients(live_only=True)
for k, gv in G.items():
    if "phi/" in k and ("fully_connected" in k) and "biases" in k:
        ref = grads[k].numpy()
        print("%.3e %s gpu=%s ref=%s" % (rel_err(gv, ref), k, np.array2string(gv, precision=5), np.array2string(ref, precision=5)))
# pre-clip mu margins
for t in range(3):
    mu = fw["mu"][t].numpy()
    print("step", t, "mu (clipped) |mu|==clip count", int((np.abs(np.abs(mu) - 0.05) < 1e-12).sum()), "of", mu.size,
          " min distance of unclipped to boundary", float(np.min(0.05 - np.abs(mu[np.abs(mu) < 0.05 - 1e-12])) if (np.abs(mu) < 0.05-1e-12).any() else -1))
