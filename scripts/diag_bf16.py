import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import seqvae_oracle as O
from gpu_util import make_pair, make_inputs, rel_err

def run(netname, dims, rng, B, **over):
    res = {}
    for operand in ("fp32", "bf16"):
        model, hp, P = make_pair(netname, dims, rng, B, operand, **over)
        x, eps = make_inputs(hp, B)
        out = model.forward(x.numpy(), None, eps.numpy(), 0.6)
        model.backward()
        res[operand] = (out, model.gradients(live_only=True), model.tc_layers)
        model.close()
    print("==", netname, dims, B, "tc_layers", res["bf16"][2], "x err", float(np.abs(res["bf16"][0]["x"] - res["fp32"][0]["x"]).max()))
    for k, g32 in res["fp32"][1].items():
        e = rel_err(res["bf16"][1][k], g32)
        flag = "  <<<<" if e > 0.1 else ""
        if "weights" in k and ("Conv" in k or "fully_connected_4" in k) or e > 0.1:
            print("  %.3e  %s%s" % (e, k, flag))

run("c_inhomog", [64, 64, 3], (-1.0, 1.0), 8, mc_steps=2)
