"""Chain noise (first slice): `add_noise_to_chain` with the fixed per-step `noise_stddevs` (reference
sequential_vae.py:233,239,1088-1091,1736; netnames c_sample_images / c_homog_sample_images :761-767).  The sample
mle_t + reg_coeff * stddev_t * N(0, I) is what chain step t + 1 reads, in training and in generation mode.

The tf.random_normal draws are INJECTED (svae_set_chain_noise_host) so that the oracle sees the same values; the in-kernel
Philox draws are checked against their distribution."""
import numpy as np
import pytest
import torch

from oracle import seqvae_oracle as O
from gpu_util import make_inputs, make_pair, oracle_mode, rel_err

pytestmark = pytest.mark.gpu

ARCH = dict(filter_sizes=[3, 8, 16, 16, 24, 24], vlae_latent_dims=[2, 3, 2, 2], mc_steps=3, noise_stddevs=[0.5, 0.25, 0.0])


def _noise(hp, B, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.randn([hp["mc_steps"], B] + hp["data_dims"], generator=g, dtype=torch.float64).float().double()


@pytest.mark.parametrize("netname", ["c_sample_images", "c_homog_sample_images"])
def test_chain_noise_forward_and_gradients_match_oracle(netname):
    B = 5
    model, hp, P = make_pair(netname, [16, 16, 3], (-1.0, 1.0), B, "fp32", **ARCH)
    assert hp["add_noise_to_chain"] and model.add_noise_to_chain
    x, eps = make_inputs(hp, B)
    ce = _noise(hp, B)
    tgt = (x * 0.9).float().double()
    model.set_chain_noise(ce.numpy())
    out = model.forward(x.numpy(), tgt.numpy(), eps.numpy(), 0.7)
    smp = model.chain_samples(B)
    model.backward()
    G = model.gradients(live_only=True)
    fw, grads = O.loss_and_grads(hp, P, x, tgt, eps, 0.7, ce)
    for t in range(hp["mc_steps"]):
        assert rel_err(out["x"][t], fw["x"][t].numpy()) < 1e-4, t
        assert rel_err(smp[t], fw["sample"][t].numpy()) < 1e-4, t
        assert abs(out["recon"][t] - float(fw["recon"][t])) < 1e-4 * float(fw["recon"][t])
    assert np.array_equal(smp[2], out["x"][2])                          # stddev 0 on the last step
    assert not np.allclose(smp[0], out["x"][0])
    assert abs(out["loss"] - float(fw["loss"])) < 1e-4 * abs(float(fw["loss"]))
    errs = [rel_err(G[k], grads[k].numpy()) for k in G if grads.get(k) is not None and float(grads[k].norm()) > 1e-9]
    assert len(errs) > 40 and np.median(errs) < 1e-4 and max(errs) < 5e-3, (np.median(errs), max(errs))
    # the injected draws matter: another set of draws moves step 1 but not step 0
    model.set_chain_noise(_noise(hp, B, seed=8).numpy())
    out2 = model.forward(x.numpy(), tgt.numpy(), eps.numpy(), 0.7)
    assert np.allclose(out2["x"][0], out["x"][0], rtol=0, atol=2e-6) and not np.allclose(out2["x"][1], out["x"][1], rtol=0, atol=1e-3)
    # training_mc_samples: the T mles followed by the T samples (:1450-1451), reg_coeff = 1
    model.set_chain_noise(ce.numpy())
    lst = model.training_mc_samples(x.numpy(), eps.numpy())
    assert len(lst) == 2 * hp["mc_steps"]
    with torch.no_grad():
        f1 = O.forward_chain(hp, P, x, x, eps, 1.0, ce)
    assert rel_err(lst[1], f1["x"][1].numpy()) < 1e-4 and rel_err(lst[hp["mc_steps"]], f1["sample"][0].numpy()) < 1e-4
    model.close()


def test_chain_noise_generation_matches_oracle():
    B = 6
    model, hp, P = make_pair("c_sample_images", [16, 16, 3], (-1.0, 1.0), B, "fp32", train=False, **ARCH)
    g = torch.Generator().manual_seed(2)
    z = torch.randn(hp["mc_steps"], B, hp["latent_dim"], generator=g, dtype=torch.float64).float().double()
    ce = _noise(hp, B)
    model.set_chain_noise(ce.numpy())
    lst = model.generate_mc_samples(None, B, z=z.numpy())
    T = hp["mc_steps"]
    assert len(lst) == 2 * T + 1                                        # generative_mles + generative_samples (x_0 first), :1424
    smp = []
    with torch.no_grad():
        mles = O.generate_chain(hp, P, z, B, ce, smp)
    for t in range(T):
        assert rel_err(lst[t], mles[t].numpy()) < 1e-4, t
        assert rel_err(lst[T + 1 + t], smp[t].numpy()) < 1e-4, t
    model.close()


@pytest.mark.parametrize("operand", ["fp32", "bf16"])
def test_chain_noise_philox_draws(operand):
    """Default draws: (sample - mle) / (reg * stddev) must be N(0, 1), reproducible for the same seed, different across steps and
    iterations; the next step's chain encoder reads the sample (bf16 family: its bf16 copy)."""
    B = 8
    model, hp, P = make_pair("c_sample_images", [32, 32, 3], (-1.0, 1.0), B, operand,
                             filter_sizes=[3, 32, 32, 64, 64, 64], vlae_latent_dims=[2, 3, 2, 2], mc_steps=3, noise_stddevs=[0.5, 0.25, 0.0])
    x, eps = make_inputs(hp, B)
    out = model.forward(x.numpy(), None, eps.numpy(), 0.6, seed=5)
    smp = model.chain_samples(B)
    n0 = (smp[0] - out["x"][0]) / (0.6 * 0.5)
    n1 = (smp[1] - out["x"][1]) / (0.6 * 0.25)
    for n in (n0, n1):
        assert abs(n.mean()) < 0.02 and abs(n.std() - 1.0) < 0.02 and abs((n ** 3).mean()) < 0.06
    assert abs(np.corrcoef(n0.ravel(), n1.ravel())[0, 1]) < 0.02
    assert np.array_equal(smp[2], out["x"][2])
    enc_in = model.block_tensor(1, "enc", 0, "input")                   # what step 1's chain encoder consumed
    assert rel_err(enc_in, smp[0]) < (1e-6 if operand == "fp32" else 4e-3)
    out_b = model.forward(x.numpy(), None, eps.numpy(), 0.6, seed=5)
    assert np.allclose(model.chain_samples(B)[0] - out_b["x"][0], smp[0] - out["x"][0], rtol=0, atol=2e-6)   # same seed, same draws
    out_c = model.forward(x.numpy(), None, eps.numpy(), 0.6, seed=6)
    assert not np.allclose(model.chain_samples(B)[0] - out_c["x"][0], smp[0] - out["x"][0], rtol=0, atol=1e-2)
    # a few train steps (the 2nd and 3rd replay the captured graph): finite, weights move
    p0 = model.get_params(live_only=True)
    losses = [model.train(x.numpy().astype(np.float32), x.numpy().astype(np.float32)) for _ in range(3)]
    assert np.all(np.isfinite(losses))
    p1 = model.get_params(live_only=True)
    assert any(not np.array_equal(p0[k], p1[k]) for k in p0)
    model.close()


def test_chain_noise_errors():
    from seqvae_b200 import _cabi
    model, hp, P = make_pair("c_inhomog", [16, 16, 3], (-1.0, 1.0), 4, "fp32", filter_sizes=[3, 8, 16, 16, 24, 24],
                             vlae_latent_dims=[2, 3, 2, 2], mc_steps=2)
    with pytest.raises(_cabi.SvaeError):
        model.set_chain_noise(np.zeros([2, 4, 16, 16, 3], np.float32))  # handle without add_noise_to_chain
    model.close()
    with pytest.raises(ValueError):
        make_pair("c_sample_images", [16, 16, 3], (-1.0, 1.0), 4, "fp32", mc_steps=3)   # noise_stddevs must have mc_steps entries
