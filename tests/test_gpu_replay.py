"""Local-replay parity of the PRODUCTION kernels inside the real chain (tests/replay.py): after one forward + backward of
the plan a train step runs, every tensor-core block's own device tensors are read back through the C ABI probe
(svae_debug_block_tensor) and the oracle layer is evaluated on them - forward contraction, batch norm + shortcut +
activation, batch-norm backward, beta / weight gradients (tc2_wgrad for the bf16 family) and input gradients, each at a
fixed bound with no chaos amplification between layers.

Bounds (norm-relative per tensor).  fp32-stored tensors: 1e-4 (fp32 accumulation over up to 4e5 products, atomics in any
order).  Tensors the bf16 family keeps ONLY as the bf16 operand copy of their consumer (activated outputs and dL/dy read
by TMA-fed kernels): 2.5e-3 = one bf16 rounding (2^-9 relative, uniform: 1.1e-3 rms) with margin - the value compared
IS the kernel's operand, so this is the storage rounding and nothing else."""
import numpy as np
import pytest

from gpu_util import make_inputs, make_pair
from replay import replay

pytestmark = pytest.mark.gpu

BOUNDS = {
    "fp32": dict(fwd_contraction=3e-5, bn_act_out=3e-5, bn_bwd_dy=1e-4, dbeta=2e-4, wgrad=1e-4, dgrad=1e-4),
    "bf16": dict(fwd_contraction=3e-5, bn_act_out=2.5e-3, bn_bwd_dy=2.5e-3, dbeta=2e-4, wgrad=1e-4, dgrad=1e-4),
}


def check(worst, operand):
    for k, bound in BOUNDS[operand].items():
        assert k in worst, "check %s never ran" % k
        assert worst[k][0] < bound, (k, worst[k], bound)


@pytest.mark.parametrize("operand", ["bf16", "fp32"])
@pytest.mark.parametrize("netname,dims,rng,B,over", [
    ("c_inhomog", [64, 64, 3], (-1.0, 1.0), 16, dict(mc_steps=3)),                # config 3 architecture
    ("m_inhomog", [32, 32, 1], (0.0, 1.0), 16, dict(mc_steps=2)),                 # config 1 architecture
    ("c_inhomog", [32, 32, 3], (0.0, 1.0), 12, dict(mc_steps=2)),                 # config 2 architecture
    ("sequential_vae_lsun", [64, 64, 3], (-1.0, 1.0), 6, dict(mc_steps=2)),       # config 4 architecture (Z = 110)
])
def test_local_replay_of_every_block(netname, dims, rng, B, over, operand):
    model, hp, P = make_pair(netname, dims, rng, B, operand, **over)
    x, eps = make_inputs(hp, B)
    tgt = (x * 0.9).float().double()
    model.forward(x.numpy(), tgt.numpy(), eps.numpy(), 0.7)
    model.backward()
    worst = replay(model, hp, P, operand)
    print("local replay %s %s: %s" % (netname, operand, {k: "%.2e @ %s" % v for k, v in worst.items()}))
    check(worst, operand)
    if operand == "bf16":
        assert model.tc_layers > 0
    model.close()


def test_probe_errors():
    """The probe refuses tensors that do not exist (no backward yet, no such block)."""
    from seqvae_b200 import _cabi

    model, hp, P = make_pair("c_inhomog", [16, 16, 3], (-1.0, 1.0), 4, "fp32",
                             filter_sizes=[3, 8, 16, 16, 24, 24], vlae_latent_dims=[2, 3, 2, 2], mc_steps=2)
    x, eps = make_inputs(hp, 4)
    with pytest.raises(_cabi.SvaeError):
        model.block_tensor(0, "enc", 0, "y")                 # before any forward
    model.forward(x.numpy(), None, eps.numpy(), 1.0)
    y = model.block_tensor(1, "enc", 0, "y")
    assert y.shape == (4, 8, 8, 8) and np.isfinite(y).all()
    with pytest.raises(_cabi.SvaeError):
        model.block_tensor(0, "enc", 0, "y")                 # step 0 has no chain encoder
    with pytest.raises(_cabi.SvaeError):
        model.block_tensor(1, "enc", 0, "dy")                # no backward yet
    with pytest.raises(_cabi.SvaeError):
        model.block_tensor(1, "tb", 9, "y")
    model.close()
