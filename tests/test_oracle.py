"""CPU tests of the oracle itself: TF-semantics pinning against an independent naive numpy restatement, finite
differences, parameter counts derived from the reference graph, golden fixtures (so that the oracle cannot drift)."""
import math
import re
from collections import OrderedDict
import os

import numpy as np
import pytest
import torch

from oracle import seqvae_oracle as O
from oracle import tf_semantics_np as TFNP

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("netname,dims,rng,ntensors,nelem,ndead", [
    ("c_inhomog", [64, 64, 3], (-1, 1), 782, 83768671, 8396800),      # SURVEY 8(c) cross-check figures
    ("m_inhomog", [32, 32, 1], (0, 1), 360, 23434437, 3280000),
    ("c_inhomog", [32, 32, 3], (0, 1), 782, 39507295, 3678208),
])
def test_param_counts_match_reference_graph(netname, dims, rng, ntensors, nelem, ndead):
    sp = O.param_specs(O.hyperparams(netname, dims, rng))
    assert len(sp) == ntensors
    assert sum(int(np.prod(s["shape"])) for s in sp) == nelem
    assert sum(int(np.prod(s["shape"])) for s in sp if s["dead"]) == ndead


def test_variable_names_follow_tf_scoping():
    sp = O.param_specs(O.hyperparams("c_inhomog", [64, 64, 3], (-1, 1)))
    names = [s["name"] for s in sp]
    assert names[0] == "phi/inference_step_0/Conv/weights"
    assert "phi/inference_step_3/fully_connected_8/biases" in names          # last sigma head (App. D)
    assert "theta/generative_encoder_step_1/BatchNorm_7/beta" in names       # enc.fc BN
    assert "theta/generative_step_0/Conv2d_transpose_6/weights" in names     # output deconv
    assert "theta/generative_step_0/Conv2d_transpose_7/weights" not in names  # no gate on step 0
    assert "theta/generative_step_5/Conv2d_transpose_7/biases" in names
    assert not any(n.startswith("theta/generative_encoder_step_0/") for n in names)


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("H", [4, 6])
def test_conv_same_matches_naive_numpy(stride, H):
    rs = np.random.RandomState(0)
    x = rs.randn(2, H, H, 3)
    w = rs.randn(4, 4, 3, 5)
    ref = TFNP.conv2d_same(x, w, stride)
    got = O.conv2d_same(torch.tensor(x), torch.tensor(w), stride).numpy()
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("stride", [1, 2])
def test_conv_transpose_is_backprop_input(stride):
    rs = np.random.RandomState(1)
    Hin = 3
    x = rs.randn(2, Hin, Hin, 5)                 # deconv input [N,H,W,Cin_t]
    w = rs.randn(4, 4, 6, 5)                     # [kh,kw,Cout_t,Cin_t]
    ref = TFNP.conv2d_backprop_input(x, w, stride, (Hin * stride, Hin * stride))
    got = O.conv2d_transpose_same(torch.tensor(x), torch.tensor(w), stride).numpy()
    assert got.shape == (2, Hin * stride, Hin * stride, 6)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("stride", [1, 2])
def test_adjoint_dot_product(stride):
    """<conv(x), y> == <x, deconv(y)>."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 8, 8, 3, generator=g, dtype=torch.float64)
    w = torch.randn(4, 4, 3, 7, generator=g, dtype=torch.float64)
    y = torch.randn(2, 8 // stride, 8 // stride, 7, generator=g, dtype=torch.float64)
    lhs = (O.conv2d_same(x, w, stride) * y).sum()
    rhs = (x * O.conv2d_transpose_same(y, w, stride)).sum()
    assert abs(float(lhs - rhs)) < 1e-10 * max(1.0, abs(float(lhs)))


def test_batch_norm_and_lrelu_semantics():
    rs = np.random.RandomState(3)
    x = rs.randn(5, 4, 4, 6) * 3 + 1
    beta = rs.randn(6)
    np.testing.assert_allclose(O.batch_norm(torch.tensor(x), torch.tensor(beta)).numpy(),
                               TFNP.batch_norm_train(x, beta), rtol=1e-12, atol=1e-12)
    v = torch.tensor([-2.0, -0.5, 0.0, 3.0], dtype=torch.float64)
    np.testing.assert_allclose(O.lrelu(v).numpy(), [-0.2, -0.05, 0.0, 3.0])   # abstract_network.py:8-10


def test_adam_matches_tf_formulation():
    hp = O.hyperparams("c_inhomog", [16, 16, 3], (-1, 1), filter_sizes=[3, 4, 4, 4, 4, 4], mc_steps=1)
    P = O.init_params(hp, 0)
    k = "phi/inference_step_0/Conv/weights"
    st = O.AdamState(P)
    g = torch.linspace(-30, 30, P[k].numel(), dtype=torch.float64).reshape(P[k].shape)
    p0 = P[k].clone().numpy()
    m = np.zeros_like(p0)
    v = np.zeros_like(p0)
    for t in (1, 2, 3):
        O.adam_apply(hp, P, {k: g}, st, 2e-4)
        p0, m, v = TFNP.adam_tf(p0, np.clip(g.numpy(), -10, 10), m, v, t, 2e-4)
    np.testing.assert_allclose(P[k].numpy(), p0, rtol=1e-12, atol=1e-15)


def _tiny():
    hp = O.hyperparams("c_inhomog", [16, 16, 3], (-1, 1), filter_sizes=[3, 4, 6, 6, 8, 8], mc_steps=2)
    g = torch.Generator().manual_seed(5)
    P = O.init_params(hp, 1)
    for k in P:       # larger weights than N(0,0.02): makes finite differences well conditioned
        if k.endswith("weights"):
            P[k] = P[k] * 5
    x = torch.rand(3, 16, 16, 3, generator=g, dtype=torch.float64) * 2 - 1
    eps = torch.randn(2, 3, hp["latent_dim"], generator=g, dtype=torch.float64)
    return hp, P, x, eps


def test_gradients_match_finite_differences():
    hp, P, x, eps = _tiny()
    fw, grads = O.loss_and_grads(hp, P, x, x, eps, 0.7)
    rs = np.random.RandomState(0)
    checked = 0
    for name in ["phi/inference_step_1/Conv_2/weights", "theta/generative_encoder_step_1/Conv/weights",
                 "theta/generative_step_0/fully_connected_4/weights", "theta/generative_step_1/Conv2d_transpose_7/weights",
                 "theta/generative_step_0/BatchNorm_6/beta", "phi/inference_step_0/fully_connected_1/biases"]:
        g = grads[name]
        assert g is not None
        idx = tuple(rs.randint(0, s) for s in g.shape)
        h = 1e-5
        vals = []
        for sgn in (+1, -1):
            Q = {k: v.clone() for k, v in P.items()}
            Q[name][idx] += sgn * h
            with torch.no_grad():
                vals.append(float(O.forward_chain(hp, Q, x, x, eps, 0.7)["loss"]))
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - float(g[idx])) <= 1e-5 * max(1.0, abs(fd)), (name, fd, float(g[idx]))
        checked += 1
    assert checked == 6


def test_inert_biases_and_dead_branch():
    """Q2/Q3: BN-shadowed biases get (numerically) zero gradient; dead-branch variables get None."""
    hp, P, x, eps = _tiny()
    _, grads = O.loss_and_grads(hp, P, x, x, eps, 1.0)
    sp = {s["name"]: s for s in O.param_specs(hp)}
    for k, g in grads.items():
        if sp[k]["dead"]:
            assert g is None
        elif sp[k]["inert"]:
            assert float(g.abs().max()) < 1e-12
        else:
            assert g is not None and torch.isfinite(g).all()


def test_gradient_flows_through_the_chain():
    """Q4: the stop_gradient at sequential_vae.py:1211-1212 is a no-op - step-0 decoder weights get gradient from the
    reconstruction loss of later steps."""
    hp, P, x, eps = _tiny()
    hp2 = dict(hp, intermediate_reconstruction=False)        # only the last step has a reconstruction term
    _, grads = O.loss_and_grads(hp2, P, x, x, eps, 0.0)
    assert float(grads["theta/generative_step_0/Conv2d_transpose/weights"].abs().max()) > 0


@pytest.mark.parametrize("case", ["tiny_c", "tiny_m", "tiny_h"])
def test_oracle_reproduces_golden(case):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    blob = np.load(os.path.join(GOLD, case + ".npz"))
    c = mg.CASES[case]
    hp = O.hyperparams(c["netname"], c["dims"], c["rng"], **c["overrides"])
    assert float(blob["preact_margin"]) > mg.MIN_MARGIN            # fixtures are sign-flip free by construction
    P = {k[2:]: torch.tensor(blob[k], dtype=torch.float64) for k in blob.files if k.startswith("P:")}
    assert list(P.keys()) == [s["name"] for s in O.param_specs(hp)]
    x, tgt, eps = (torch.tensor(blob[k]) for k in ("x", "tgt", "eps"))
    fw, grads = O.loss_and_grads(hp, P, x, tgt, eps, float(blob["reg"]))
    np.testing.assert_allclose(torch.stack(fw["mu"]).numpy(), blob["mu"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(torch.stack(fw["sigma"]).numpy(), blob["sigma"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(torch.stack(fw["x"]).numpy(), blob["xs"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(float(fw["loss"]), float(blob["loss"]), rtol=1e-10)
    for k, g in grads.items():
        if g is None:
            assert "G:" + k not in blob.files
        else:
            gold = blob["G:" + k]
            np.testing.assert_allclose(g.numpy(), gold, rtol=1e-5, atol=1e-7 * max(1.0, float(np.abs(gold).max())))
    with torch.no_grad():
        gen = O.generate_chain(hp, P, torch.tensor(blob["z"]), c["B"])
    np.testing.assert_allclose(torch.stack(gen).numpy(), blob["gen"], rtol=1e-9, atol=1e-11)


def test_train_wrapper_schedules():
    """train(): iteration / lr decay / reg_coeff warm-up and the final_loss/H/W return (sequential_vae.py:1351-1375)."""
    hp = O.hyperparams("c_inhomog", [16, 16, 3], (-1, 1), filter_sizes=[3, 4, 4, 4, 4, 4], mc_steps=2)
    m = O.OracleModel(hp, seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 16, 16, 3, generator=g, dtype=torch.float64) * 2 - 1
    eps = torch.randn(2, 2, hp["latent_dim"], generator=g, dtype=torch.float64)
    r, fw, _ = m.train(x, x, eps)
    assert m.iteration == 1 and m.adam.t == 1
    assert math.isclose(r, float(fw["final_loss"]) / 16 / 16)
    reg = 1 - math.exp(-1 / 5000.0)
    total = sum(16 * float(a) + reg * float(b) for a, b in zip(fw["recon"], fw["kl"]))
    assert math.isclose(float(fw["loss"]), total, rel_tol=1e-12)


def test_shared_scopes_gradient_is_the_sum_over_steps():
    """Homogeneous chain (share_theta_weights / share_phi_weights, sequential_vae.py:213-214): evaluating the SAME values
    through per-step scopes gives the same forward, and the gradient of a shared variable is the sum of the per-step
    gradients - what TF's tf.AUTO_REUSE variable sharing computes and what libsvae's tied slices + tie_reduce restate."""
    over = dict(filter_sizes=[3, 8, 16, 16, 24, 24], vlae_latent_dims=[2, 3, 2, 2], mc_steps=3)
    hp_h = O.hyperparams("sequential_vae_celebA_homog", [16, 16, 3], (-1.0, 1.0), **over)
    hp_i = O.hyperparams("c_inhomog", [16, 16, 3], (-1.0, 1.0), **over)
    assert hp_h["share_theta_weights"] and hp_h["share_phi_weights"] and not hp_i["share_theta_weights"]
    P_h = O.init_params(hp_h, 3)
    g = torch.Generator().manual_seed(5)
    for k in P_h:
        if k.endswith("/beta") or k.endswith("/biases"):
            P_h[k] = 0.1 * torch.randn(P_h[k].shape, generator=g, dtype=torch.float64)

    def shared_name(k):
        m = re.match(r"(phi/inference|theta/generative_encoder|theta/generative)_step_(\d+)/(.*)", k)
        scope, t, rest = m.group(1), int(m.group(2)), m.group(3)
        if scope == "theta/generative" and t == 0:
            return k
        return scope + "_network/" + rest

    P_i = OrderedDict((s["name"], P_h[shared_name(s["name"])].clone()) for s in O.param_specs(hp_i))
    x = torch.rand(4, 16, 16, 3, generator=g, dtype=torch.float64) * 2 - 1
    eps = torch.randn(3, 4, hp_h["latent_dim"], generator=g, dtype=torch.float64)
    fw_h, g_h = O.loss_and_grads(hp_h, P_h, x, x, eps, 0.7)
    fw_i, g_i = O.loss_and_grads(hp_i, P_i, x, x, eps, 0.7)
    assert abs(float(fw_h["loss"]) - float(fw_i["loss"])) < 1e-12
    for t in range(3):
        assert torch.equal(fw_h["x"][t], fw_i["x"][t])
    summed = {}
    for k, v in g_i.items():
        if v is not None:
            summed[shared_name(k)] = summed.get(shared_name(k), 0) + v
    n_shared = 0
    for k, v in g_h.items():
        if v is None:
            assert k not in summed
            continue
        assert torch.allclose(v, summed[k], rtol=1e-10, atol=1e-14), k
        n_shared += k.startswith(("phi/inference_network", "theta/generative_network", "theta/generative_encoder_network"))
    assert n_shared > 40
    # the recognition net is one function of x for every step: identical mu_t, sigma_t along the chain
    assert torch.equal(fw_h["mu"][0], fw_h["mu"][2]) and torch.equal(fw_h["sigma"][0], fw_h["sigma"][1])


def test_apply_noise_restatement():
    """oracle.apply_noise restates trainer.py:69-78 given the three random fields."""
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, size=(4, 8, 8, 3))
    keep = rng.binomial(1, 0.9, size=x.shape)
    salt = rng.binomial(1, 0.1, size=x.shape)
    gauss = rng.normal(scale=0.1, size=x.shape)
    out = O.apply_noise(x, keep, salt, gauss, (-1.0, 1.0))
    assert out.min() >= -1 and out.max() <= 1
    i = (keep == 1) & (salt == 0)
    np.testing.assert_allclose(out[i], np.clip(x[i] + gauss[i], -1, 1))
    i = (keep == 0) & (salt == 1)
    np.testing.assert_allclose(out[i], np.clip(1 + gauss[i], -1, 1))
    i = (keep == 0) & (salt == 0)
    np.testing.assert_allclose(out[i], gauss[i])
    assert O.NOISE_DEFAULTS == dict(pepper_prob=0.1, salt_prob=0.1, gaussian_noise_scale=0.1)


def test_chain_noise_semantics_and_gradients():
    """add_noise_to_chain (sequential_vae.py:1088-1091, netname c_sample_images :761): the sample mle_t + reg * stddev_t * eps_t is
    what step t + 1 reads; the losses stay on the mles; the gradient flows through the additive noise unchanged."""
    hp, P, x, eps = _tiny()
    hp = dict(hp, add_noise_to_chain=True, noise_stddevs=[0.5, 0.0])
    g = torch.Generator().manual_seed(3)
    ce = torch.randn(2, *x.shape, generator=g, dtype=torch.float64)
    base = O.forward_chain(dict(hp, add_noise_to_chain=False), P, x, x, eps, 0.7)
    with torch.no_grad():
        zero = O.forward_chain(hp, P, x, x, eps, 0.7, torch.zeros_like(ce))
        fw = O.forward_chain(hp, P, x, x, eps, 0.7, ce)
    assert all(torch.equal(a, b) for a, b in zip(zero["x"], base["x"])) and float(zero["loss"]) == float(base["loss"])
    assert torch.allclose(fw["sample"][0], fw["x"][0] + 0.7 * 0.5 * ce[0], rtol=0, atol=1e-15)
    assert torch.equal(fw["sample"][1], fw["x"][1])                     # stddev 0 on the last step (:152-153)
    assert torch.equal(fw["x"][0], base["x"][0]) and not torch.allclose(fw["x"][1], base["x"][1])   # step 1 reads the sample
    _, grads = O.loss_and_grads(hp, P, x, x, eps, 0.7, ce)
    name, h = "theta/generative_step_0/Conv2d_transpose_6/weights", 1e-5    # reaches the loss of step 1 through the sample
    idx = (1, 2, 0, 3)
    vals = []
    for sgn in (+1, -1):
        Q = {k: v.clone() for k, v in P.items()}
        Q[name][idx] += sgn * h
        with torch.no_grad():
            vals.append(float(O.forward_chain(hp, Q, x, x, eps, 0.7, ce)["loss"]))
    fd = (vals[0] - vals[1]) / (2 * h)
    assert abs(fd - float(grads[name][idx])) <= 1e-5 * max(1.0, abs(fd))
    # generation: reg_coeff is the placeholder's default 1 (:917), the samples feed the next step
    z = torch.randn(2, 3, hp["latent_dim"], generator=g, dtype=torch.float64)
    smp = []
    with torch.no_grad():
        mles = O.generate_chain(hp, P, z, 3, ce, smp)
        clean = O.generate_chain(dict(hp, add_noise_to_chain=False), P, z, 3)
    assert torch.allclose(smp[0], mles[0] + 0.5 * ce[0], rtol=0, atol=1e-15) and torch.equal(mles[0], clean[0])
    assert not torch.allclose(mles[1], clean[1])
    # the netname rows: same variables as the noise-free nets, the reference's default stddevs (:239)
    hs = O.hyperparams("c_sample_images", [64, 64, 3], (-1, 1))
    assert hs["add_noise_to_chain"] and hs["noise_stddevs"] == [0.5 ** k for k in range(1, 8)] + [0]
    assert [s["name"] for s in O.param_specs(hs)] == [s["name"] for s in O.param_specs(O.hyperparams("c_inhomog", [64, 64, 3], (-1, 1)))]
    hh = O.hyperparams("c_homog_sample_images", [64, 64, 3], (-1, 1))
    assert hh["share_theta_weights"] and hh["share_phi_weights"] and hh["add_noise_to_chain"]
