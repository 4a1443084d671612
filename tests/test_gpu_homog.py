"""GPU parity of the homogeneous (weight-shared) chains - share_theta_weights / share_phi_weights, reference
sequential_vae.py:213-214,1573-1577,1683-1687,1757-1761; netnames sequential_vae_celebA_homog (:675), c_homog (:730),
sequential_vae_celebA_homog_fixed_length (:709) - through the same Python surface + C ABI as tests/test_gpu_chain.py and
at the same bounds: per-step mu, sigma, x_t, ELBO terms, the gradient of every shared variable (sum over the chain),
clipped-Adam trajectories and generation vs the fp64 oracle on identical weights / inputs / injected eps."""
import math

import numpy as np
import pytest
import torch

import seqvae_b200 as S
from oracle import seqvae_oracle as O
from gpu_util import TINY, make_inputs, make_pair, oracle_mode, rel_err
from test_gpu_chain import _check_forward, _check_grads, _oracle_pair

pytestmark = pytest.mark.gpu
# bf16 family: the kernels are compared with the oracle evaluating the SAME graph with the same operands rounded to bf16.  An
# activation that sits on a bf16 rounding boundary can round the other way in the kernel (fp32 accumulation order) than in the
# emulation (fp64 accumulation): that element then differs by one bf16 ulp, 2^-8 relative, and so does what it feeds.  The
# forward bound for the bf16 family is therefore one bf16 ulp of an O(1) activation at step 0, growing 3.5x per chain step (the
# amplification of any perturbation by this chain at random init, measured on the CPU oracle: tests/test_gpu_chain.py
# header); the fp32 family keeps a flat 1e-3.
FWD_TOL = {"fp32": 1e-3, "bf16": 2.0 ** -8}
FWD_GROWTH = {"fp32": 1.0, "bf16": 3.5}
SHARED_SCOPES = ("phi/inference_network/", "theta/generative_encoder_network/", "theta/generative_network/")


def _arena(model, which="param"):
    return model.read_arena(which)


def _slices_identical(model, arena):
    """Every per-step slice of a shared variable must be bit-identical to the first one."""
    shared = 0
    for p in model.param_table:
        offs = model.param_slices(p["name"])
        assert offs[0] == p["offset"]
        if p["name"].startswith(SHARED_SCOPES):
            assert len(offs) >= 2, p["name"]
            shared += 1
        else:
            assert len(offs) == 1, p["name"]
        first = arena[offs[0]:offs[0] + p["numel"]]
        for o in offs[1:]:
            assert np.array_equal(first, arena[o:o + p["numel"]]), p["name"]
    return shared


@pytest.mark.parametrize("operand", ["fp32", "bf16"])
@pytest.mark.parametrize("netname,dims,rng,B,over", [
    ("sequential_vae_celebA_homog", [16, 16, 3], (-1.0, 1.0), 5, dict(TINY, mc_steps=3)),
    ("sequential_vae_celebA_homog_fixed_length", [16, 16, 3], (-1.0, 1.0), 8, dict(TINY, mc_steps=3)),   # theta only
    ("c_homog_v1", [32, 32, 3], (0.0, 1.0), 8, dict(mc_steps=3)),                 # narrow filters [3,16,32,64,128,384], Z=48
    ("sequential_vae_celebA_homog", [64, 64, 3], (-1.0, 1.0), 6, dict(mc_steps=3)),   # benchmarked architecture, shared
])
def test_homog_forward_and_gradients_match_oracle(netname, dims, rng, B, over, operand):
    model, hp, P = make_pair(netname, dims, rng, B, operand, **over)
    assert [p["name"] for p in model.param_table] == [s["name"] for s in O.param_specs(hp)]
    x, eps = make_inputs(hp, B)
    tgt = (x * 0.9).float().double()
    with oracle_mode(operand):
        fw, grads, fw32, g32 = _oracle_pair(hp, P, x, tgt, eps, 0.6)
    out = model.forward(x.numpy(), tgt.numpy(), eps.numpy(), 0.6)
    _check_forward(out, fw, operand, fw32, tol=FWD_TOL[operand], growth=FWD_GROWTH[operand])
    model.backward()
    _check_grads(model, grads, hp, operand, g32)
    # every slice of a shared variable holds the summed gradient
    assert _slices_identical(model, _arena(model, "grad")) > 40
    model.close()


def test_homog_gradient_equals_sum_of_untied_steps():
    """Same values evaluated through an inhomogeneous handle: the shared variable's gradient is the sum of the per-step
    gradients (fp32 kernels, same kernels in both handles, so the two sides differ only in fp32 summation order)."""
    B, over = 5, dict(TINY, mc_steps=3)
    hom, hp, P = make_pair("sequential_vae_celebA_homog", [16, 16, 3], (-1.0, 1.0), B, "fp32", **over)
    ds = S.SyntheticDataset("x", B, data_dims=[16, 16, 3], data_range=[-1.0, 1.0])
    inh = S.SequentialVAE(ds, B, "c_inhomog", operand_dtype="fp32", restore=False, **over)

    def shared_name(k):
        import re
        m = re.match(r"(phi/inference|theta/generative_encoder|theta/generative)_step_(\d+)/(.*)", k)
        scope, t, rest = m.group(1), int(m.group(2)), m.group(3)
        return k if (scope == "theta/generative" and t == 0) else scope + "_network/" + rest

    inh.set_params({p["name"]: P[shared_name(p["name"])].numpy() for p in inh.param_table})
    x, eps = make_inputs(hp, B)
    a = hom.forward(x.numpy(), None, eps.numpy(), 0.5)
    b = inh.forward(x.numpy(), None, eps.numpy(), 0.5)
    np.testing.assert_allclose(a["x"], b["x"], rtol=0, atol=1e-5)
    hom.backward()
    inh.backward()
    Gh, Gi = hom.gradients(live_only=True), inh.gradients(live_only=True)
    summed = {}
    for k, v in Gi.items():
        summed[shared_name(k)] = summed.get(shared_name(k), 0) + v.astype(np.float64)
    gmax = max(float(np.abs(v).max()) for v in Gh.values())
    errs = {}
    for k, v in Gh.items():
        ref = summed[k]
        errs[k] = float(np.linalg.norm(v - ref) / max(np.linalg.norm(ref), 1e-6 * gmax * math.sqrt(ref.size)))
    # the two handles sum their batch-norm statistics with atomics: a unit within rounding of zero may flip in one of them
    assert float(np.percentile(list(errs.values()), 90)) < 5e-4, max(errs.items(), key=lambda kv: kv[1])
    assert max(errs.values()) < 0.1, max(errs.items(), key=lambda kv: kv[1])
    hom.close()
    inh.close()


@pytest.mark.parametrize("operand", ["fp32", "bf16"])
def test_homog_train_trajectory_keeps_slices_tied(operand):
    """Clipped-Adam steps through train() (the first one eager, the following ones as the captured CUDA graph): every live
    parameter follows the oracle's trajectory and the per-step slices of each shared variable stay bit-identical."""
    B = 4
    model, hp, P = make_pair("sequential_vae_celebA_homog", [16, 16, 3], (-1.0, 1.0), B, operand, **dict(TINY, mc_steps=3))
    om = O.OracleModel(hp, seed=0)
    om.P = {k: v.clone() for k, v in P.items()}
    om.adam = O.AdamState(om.P)
    assert _slices_identical(model, _arena(model)) > 40
    for it in range(4):
        x, eps = make_inputs(hp, B, seed=10 + it)
        with oracle_mode(operand):
            r_ref, fw, _ = om.train(x, x, eps, update_inert=False)
        r = model.train(x.numpy().astype(np.float32), x.numpy().astype(np.float32), eps.numpy())
        tol = 2e-3 if operand == "fp32" else 0.1
        assert math.isclose(r, r_ref, rel_tol=tol), (it, r, r_ref)
        assert math.isclose(model.last_losses["loss"], float(fw["loss"]), rel_tol=tol)
    _slices_identical(model, _arena(model))
    if operand == "fp32":
        got = model.get_params(live_only=True)
        for k, v in got.items():
            upd_ref = om.P[k].numpy() - P[k].numpy()
            upd = v.astype(np.float64) - P[k].numpy()
            if np.linalg.norm(upd_ref) > 1e-7:
                assert rel_err(upd, upd_ref) < 5e-2, k
    model.close()


@pytest.mark.parametrize("train", [True, False])
def test_homog_generation_matches_oracle(train):
    """A homogeneous chain generates with one decoder for all steps >= 1: chains longer than the trained one are the use
    (sequential_vae.py:730-733, mc_steps 25)."""
    B = 6
    over = dict(TINY, mc_steps=6)
    model, hp, P = make_pair("c_homog", [16, 16, 3], (-1.0, 1.0), B, "fp32", train=train, **over)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(6, B, hp["latent_dim"], generator=g, dtype=torch.float64).float().double()
    with torch.no_grad():
        ref = O.generate_chain(hp, P, z, B)
    gen = model.generate_mc_samples(None, B, z=z.numpy())
    # the chain at random init amplifies fp32 rounding ~3.5x per step (tests/test_gpu_chain.py): bound grows with the step
    for t in range(6):
        np.testing.assert_allclose(gen[1 + t], ref[t].numpy(), rtol=0, atol=1e-3 * 3.5 ** max(0, t - 2))
    model.close()


def test_homog_checkpoint_round_trip(tmp_path):
    """save_network / load_network (abstract_network.py:124-152) with shared scopes: one entry per TF variable, restored into
    every slice, Adam slots included."""
    B = 4
    over = dict(TINY, mc_steps=3)
    ds = S.SyntheticDataset("x", B, data_dims=[16, 16, 3], data_range=[-1.0, 1.0])
    a = S.SequentialVAE(ds, B, "sequential_vae_celebA_homog", base_dir=str(tmp_path / "m"), operand_dtype="fp32", restore=False, **over)
    x = ds.next_batch(B)
    for _ in range(2):
        a.train(x, x)
    path = a.save_network()
    blob = np.load(path)
    names = [p["name"] for p in a.param_table]
    live = [p["name"] for p in a.param_table if not p["flags"] & S._cabi.PF_DEAD]
    assert all(n in blob.files for n in names)
    assert all(n + "/Adam" in blob.files and n + "/Adam_1" in blob.files for n in live)
    assert "theta/generative_network/BatchNorm/moving_mean" in blob.files and "beta1_power" in blob.files
    assert not any("_step_1/" in n or "_step_2/" in n for n in blob.files)
    b = S.SequentialVAE(ds, B, "sequential_vae_celebA_homog", base_dir=str(tmp_path / "m"), operand_dtype="fp32", restore=True, **over)
    assert b.iteration == 2
    assert np.array_equal(_arena(a), _arena(b))
    eps = np.random.default_rng(0).normal(size=(3, B, a.latent_dim)).astype(np.float32)
    ra, rb = a.train(x, x, eps), b.train(x, x, eps)
    assert math.isclose(ra, rb, rel_tol=2e-3)     # two runs from bit-identical state: the atomics' order, and rarely one activation at the noise level
    _slices_identical(b, _arena(b))
    a.close()
    b.close()


def test_c_homog_full_chain_length_properties():
    """c_homog at its real chain length (T = 25, sequential_vae.py:730-733) on the benchmarked architecture, bf16 family, small
    batch: finite per-step ELBO terms, total consistent with them, every live variable updated, slices tied, and the
    variable count independent of T."""
    ds = S.SyntheticDataset("celebA", 8)
    model = S.SequentialVAE(ds, 8, "c_homog", operand_dtype="bf16", restore=False)
    assert model.mc_steps == 25 and len(model.param_table) == 136
    x = ds.next_batch(8)
    before = model.get_params(live_only=True)
    for it in range(4):
        model.train(x, x)
        ls = model.last_losses
        assert all(np.isfinite(ls["recon"])) and all(np.isfinite(ls["kl"])) and len(ls["recon"]) == 25
        reg = 1 - math.exp(-(it + 1) / 5000.0)
        # the KL term only enters on steps 0..7: regularized_steps is range(8), fixed before the netname row sets T = 25
        # (sequential_vae.py:224 vs :733)
        total = sum(16 * r for r in ls["recon"]) + reg * sum(ls["kl"][:8])
        assert math.isclose(ls["loss"], total, rel_tol=1e-4)
    after = model.get_params(live_only=True)
    moved = [k for k in before if not np.array_equal(before[k], after[k])]
    assert len(moved) == len(before), sorted(set(before) - set(moved))[:5]
    assert all(np.isfinite(v).all() for v in after.values())
    assert _slices_identical(model, _arena(model)) > 40
    model.close()


def test_checkpoint_resume_equals_uninterrupted_training(tmp_path):
    """Inhomogeneous chain: train 2 steps, save, restore into a fresh handle, train 1 more == 3 uninterrupted steps (weights,
    both Adam slots, step count, schedules); the file holds the TF Saver's variable set (checkpoint.tf_checkpoint_layout)."""
    from seqvae_b200.checkpoint import tf_checkpoint_layout

    B = 4
    ds = S.SyntheticDataset("x", B, data_dims=[16, 16, 3], data_range=[-1.0, 1.0])
    rng = np.random.default_rng(0)
    xs = [ds.next_batch(B) for _ in range(3)]
    es = [rng.normal(size=(2, B, 9)).astype(np.float32) for _ in range(3)]
    a = S.SequentialVAE(ds, B, "c_inhomog", base_dir=str(tmp_path / "a"), operand_dtype="fp32", restore=False, **TINY)
    c = S.SequentialVAE(ds, B, "c_inhomog", base_dir=str(tmp_path / "c"), operand_dtype="fp32", restore=False, **TINY)
    c.set_params(a.get_params())
    for i in range(2):
        a.train(xs[i], xs[i], es[i])
    path = a.save_network()
    blob = np.load(path)
    want = [e[0] for e in tf_checkpoint_layout(a.param_table, True)]
    assert sorted(want + ["__adam_t", "__iteration", "__learning_rate"]) == sorted(blob.files)
    assert abs(float(blob["beta1_power"]) - 0.9 ** 3) < 1e-6 and not blob["phi/inference_step_0/BatchNorm/moving_mean"].any()
    b = S.SequentialVAE(ds, B, "c_inhomog", base_dir=str(tmp_path / "a"), operand_dtype="fp32", restore=True, **TINY)
    assert b.iteration == 2
    for w in ("param", "adam_m", "adam_v"):
        assert np.array_equal(a.read_arena(w), b.read_arena(w)), w
    rb = b.train(xs[2], xs[2], es[2])
    for i in range(3):
        rc = c.train(xs[i], xs[i], es[i])
    assert math.isclose(rb, rc, rel_tol=5e-3)     # c's first two steps were a separate run: a few Adam updates may have gone the other way
    pb, pc = b.get_params(live_only=True), c.get_params(live_only=True)
    n = bad = 0
    for k in pb:
        d = np.abs(pb[k].astype(np.float64) - pc[k])
        assert d.max() <= 2 * 3 * 2e-4 * 1.01, k
        n += d.size
        bad += int((d > 2e-5).sum())
    assert bad <= 1e-3 * n, (bad, n)
    # a second save moves the first file to <models>/old like the reference (abstract_network.py:131-133)
    a.save_network()
    assert (tmp_path / "old" / "c_inhomog_v0.npz").exists()
    for m in (a, b, c):
        m.close()


@pytest.mark.parametrize("netname,dims,rng,B,over", [
    ("vlae_celebA", [16, 16, 3], (-1.0, 1.0), 6, dict(filter_sizes=TINY["filter_sizes"])),        # :704-707: mc_steps 1, Z = 64
    ("c_homog_one_step", [32, 32, 3], (0.0, 1.0), 8, {}),                                          # :281-288: shared scopes, one step
])
def test_single_step_netnames(netname, dims, rng, B, over):
    """mc_steps = 1 (a plain VLAE): no chain encoder, no gate, nothing to tie - forward, gradients, one Adam step, generation."""
    model, hp, P = make_pair(netname, dims, rng, B, "fp32", **over)
    assert model.mc_steps == 1 and [p["name"] for p in model.param_table] == [s["name"] for s in O.param_specs(hp)]
    assert all(len(model.param_slices(p["name"])) == 1 for p in model.param_table)
    x, eps = make_inputs(hp, B)
    fw, grads, fw32, g32 = _oracle_pair(hp, P, x, x, eps, 0.7)
    out = model.forward(x.numpy(), None, eps.numpy(), 0.7)
    _check_forward(out, fw, "fp32", fw32)
    model.backward()
    _check_grads(model, grads, hp, "fp32", g32)
    r = model.train(x.numpy().astype(np.float32), x.numpy().astype(np.float32), eps.numpy())
    assert np.isfinite(r)
    r2 = model.train(x.numpy().astype(np.float32), x.numpy().astype(np.float32), eps.numpy())       # graph replay
    assert np.isfinite(r2) and r2 != r
    g = torch.Generator().manual_seed(3)
    z = torch.randn(1, B, hp["latent_dim"], generator=g, dtype=torch.float64).float().double()
    P2 = {k: torch.tensor(v, dtype=torch.float64) for k, v in model.get_params().items()}
    with torch.no_grad():
        ref = O.generate_chain(hp, P2, z, B)
    gen = model.generate_mc_samples(None, B, z=z.numpy())
    assert len(gen) == 2
    np.testing.assert_allclose(gen[1], ref[0].numpy(), rtol=1e-3, atol=1e-3)
    model.close()
