"""BASELINE.json's configurations at their FULL sizes.  The benchmarked configuration (config 3: CelebA-64 c_inhomog,
B=100, T=8) is compared with the oracle AT FULL SIZE (test_full_size_oracle_parity: the fp64 oracle takes ~10 s per
forward + backward on the box's host cores) and its production kernels are replayed block by block
(test_full_size_local_replay_bf16).  The other configurations go through properties that do not need an oracle run (the same
architectures are checked against the oracle at small batch in tests/test_gpu_chain.py and tests/test_gpu_replay.py): config 1 MNIST m_inhomog B=100, config 2 CIFAR-shaped c_inhomog B=100, config 4 LSUN long chain
(mc_steps 25, B=256), config 5 generation B=4096.  Config 3 (CelebA B=100) is test_full_size_properties_celeba_b100.

Properties: finite per-step ELBO terms in the range expected at initialisation; the returned total equals
sum_t 16*recon_t + reg*KL_t (sequential_vae.py:1168-1176); sigma in (0,1) (:1594); reconstructions inside dataset.range
(:1721); dead / inert variables untouched and every live variable updated by a step (Q2, Q3); the KL of a fresh net is small
and positive; generation is a function of (z, batch statistics) only: deterministic for a fixed z, and permuting the batch
of z permutes the samples (batch-norm statistics are permutation invariant)."""
import math

import numpy as np
import pytest

import seqvae_b200 as S
from seqvae_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dataset,netname,B,operand,over", [
    ("mnist", "m_inhomog", 100, "fp32", {}),                                  # config 1 (strict-parity kernel family)
    ("mnist", "m_inhomog", 100, "bf16", {}),
    ("cifar", "c_inhomog", 100, "bf16", {}),                                  # config 2
    ("lsun", "sequential_vae_lsun", 256, "bf16", dict(mc_steps=25)),          # config 4, the longest chain the reference uses (:733)
])
def test_full_size_train_properties(dataset, netname, B, operand, over):
    ds = S.SyntheticDataset(dataset, B)
    model = S.SequentialVAE(ds, B, netname, operand_dtype=operand, restore=False, **over)
    T = model.mc_steps
    x = ds.next_batch(B)
    lo, hi = ds.range
    out = model.forward(x, None, None, 1.0, seed=1)
    assert out["mu"].shape == (T, B, model.latent_dim) and np.isfinite(out["mu"]).all()
    assert (out["sigma"] > 0).all() and (out["sigma"] < 1).all()
    assert out["x"].min() >= lo - 1e-6 and out["x"].max() <= hi + 1e-6
    # an untrained net reconstructs U[lo,hi] data no better than its mean: recon ~ var + bias^2, O((hi-lo)^2)
    assert all(0.02 * (hi - lo) ** 2 < r < (hi - lo) ** 2 for r in out["recon"]), out["recon"]
    assert all(0 < k < 5 for k in out["kl"]), out["kl"]
    before = model.get_params()
    for it in range(3):
        model.train(x, x)
        ls = model.last_losses
        assert all(np.isfinite(ls["recon"])) and all(np.isfinite(ls["kl"])) and len(ls["recon"]) == T
        reg = 1 - math.exp(-(it + 1) / 5000.0)
        total = sum(16 * r + reg * k for r, k in zip(ls["recon"], ls["kl"]))
        assert math.isclose(ls["loss"], total, rel_tol=1e-4)
    after = model.get_params()
    for p in model.param_table:
        same = np.array_equal(before[p["name"]], after[p["name"]])
        if p["flags"] & (_cabi.PF_DEAD | _cabi.PF_INERT):
            assert same, p["name"]
        else:
            assert not same, p["name"]
            assert np.isfinite(after[p["name"]]).all(), p["name"]
    model.close()


def test_full_size_generation_properties():
    """Config 5: generation-only chain, CelebA-64, B=4096, all 8 steps, forward-only handle."""
    B = 4096
    ds = S.SyntheticDataset("celebA", B)
    model = S.SequentialVAE(ds, B, "c_inhomog", operand_dtype="bf16", train=False, restore=False)
    T, Z = model.mc_steps, model.latent_dim
    z = np.random.default_rng(0).normal(size=(T, B, Z)).astype(np.float32)
    a = np.stack(model.generate_mc_samples(None, B, z=z)[1:])
    assert a.shape == (T, B, 64, 64, 3) and np.isfinite(a).all()
    assert a.min() >= -1.0 - 1e-6 and a.max() <= 1.0 + 1e-6
    b = np.stack(model.generate_mc_samples(None, B, z=z)[1:])
    # same z => same chain up to the summation order of the batch statistics (atomics): tight at the first step, and -
    # because the chain at random init amplifies any perturbation step by step (tests/test_gpu_chain.py) - loose at the last
    assert np.abs(a[0] - b[0]).max() < 2e-2 and np.abs(a[0] - b[0]).mean() < 1e-4 and np.abs(a[-1] - b[-1]).mean() < 0.05
    del b
    perm = np.random.default_rng(1).permutation(B)
    c = np.stack(model.generate_mc_samples(None, B, z=np.ascontiguousarray(z[:, perm]))[1:])
    assert np.abs(c[0] - a[0][perm]).max() < 2e-2 and np.abs(c[0] - a[0][perm]).mean() < 1e-4
    assert np.abs(c[-1] - a[-1][perm]).mean() < 0.05
    del c
    # different latents give different samples, and samples within a batch differ from each other
    assert np.abs(a[-1][0] - a[-1][1]).max() > 1e-3
    d = np.stack(model.generate_mc_samples(None, B, seed=9)[1:])       # device-side Philox z
    assert np.isfinite(d).all() and np.abs(d[0] - a[0]).max() > 1e-3
    model.close()


def _tensor_stats(G, grads, sp):
    """per-tensor norm-relative error and cosine of the live gradient tensors"""
    rel, cos = {}, {}
    for k, ref in grads.items():
        if ref is None or sp[k]["inert"]:
            continue
        r = ref.double().numpy().ravel()
        g = np.asarray(G[k], np.float64).ravel()
        nr, ng = np.linalg.norm(r), np.linalg.norm(g)
        if nr < 1e-12:
            continue
        rel[k] = float(np.linalg.norm(g - r) / nr)
        cos[k] = float(g @ r / (nr * ng + 1e-300))
    return rel, cos


def _distance(out_x, out_mu, out_sd, out_recon, out_kl, out_loss, G, fw, grads, sp, T):
    """distance of one evaluation (device or CPU) from the fp64 oracle: forward tensors per step, loss, gradient tensors"""
    fwd = {}
    for key, val in (("mu", out_mu), ("sigma", out_sd), ("x", out_x)):
        fwd[key] = [float(np.abs(np.asarray(val[t], np.float64) - fw[key][t].numpy()).max() / max(1.0, float(fw[key][t].abs().max())))
                    for t in range(T)]
    for key, val in (("recon", out_recon), ("kl", out_kl)):
        fwd[key] = [abs(float(val[t]) - float(fw[key][t])) / max(abs(float(fw[key][t])), 0.1) for t in range(T)]
    fwd["loss"] = abs(float(out_loss) - float(fw["loss"])) / abs(float(fw["loss"]))
    rel, cos = _tensor_stats(G, grads, sp)
    vals, cvals = np.array(list(rel.values())), np.array(list(cos.values()))
    worst = max(rel.items(), key=lambda kv: kv[1])
    return dict(forward_max_rel_err_per_step=fwd, grad_tensors=len(vals),
                grad_rel_err=dict(median=float(np.median(vals)), p95=float(np.percentile(vals, 95)), max=float(vals.max()),
                                  worst_tensor=worst[0]),
                grad_cosine=dict(median=float(np.median(cvals)), p05=float(np.percentile(cvals, 5)), min=float(cvals.min())),
                frac_tensors_within_1e_3=float((vals < 1e-3).mean()))


def test_full_size_oracle_parity():
    """The benchmarked configuration at its full size (CelebA-64 c_inhomog, B=100, T=8) against the fp64 oracle, same weights,
    inputs and injected eps.  Three evaluations are placed at their measured distance from the fp64 oracle and recorded in
    gpurun_out/parity_fullsize.json (bench.py quotes the committed copy as `parity`):
      * the fp32 CPU oracle itself (plain fp32 PyTorch arithmetic on the same graph) - what ANY fp32 evaluation can reach:
        this chain at random init amplifies rounding noise ~2.4x per chain step forward and more in the backward, so the
        1e-3 north-star bound holds for the forward tensors of the first steps and is out of reach of fp32 arithmetic for the
        T=8 gradients (measured here, not assumed);
      * the fp32 kernel family: must be no further from fp64 than 2x the fp32 CPU oracle (or inside 1e-3 where that is
        reachable) on every quantity - i.e. it is as good as fp32 arithmetic gets;
      * the bf16 tensor-core family against the UNROUNDED fp64 oracle: measured and recorded; gradient direction must be kept
        (cosine).  Its kernels are held to tight bounds by the local replay (test_full_size_local_replay_bf16)."""
    import json
    import os
    import time

    import torch

    from oracle import seqvae_oracle as O
    from gpu_util import make_inputs, make_pair

    B = 100
    report = {"config": "celeba64 c_inhomog B=100 T=8, reference initialisers (seed 0) with perturbed betas/biases, "
                        "x ~ U[-1,1], target = 0.9 x, injected eps, reg_coeff 0.6; distances from the fp64 CPU oracle"}
    fw = grads = cpu32 = None
    for operand in ("fp32", "bf16"):
        model, hp, P = make_pair("c_inhomog", [64, 64, 3], (-1.0, 1.0), B, operand)
        x, eps = make_inputs(hp, B)
        tgt = (x * 0.9).float().double()
        T = hp["mc_steps"]
        sp = {s_["name"]: s_ for s_ in O.param_specs(hp)}
        if fw is None:
            torch.set_num_threads(os.cpu_count() or 1)
            t0 = time.time()
            fw, grads = O.loss_and_grads(hp, P, x, tgt, eps, 0.6)
            report["oracle_fp64_seconds"] = round(time.time() - t0, 1)
            f32, g32 = O.loss_and_grads(hp, {k: v.float() for k, v in P.items()}, x.float(), tgt.float(), eps.float(), 0.6)
            cpu32 = _distance(f32["x"], f32["mu"], f32["sigma"], f32["recon"], f32["kl"], f32["loss"],
                              {k: v.numpy() for k, v in g32.items() if v is not None}, fw, grads, sp, T)
            report["fp32_cpu_oracle"] = cpu32
        out = model.forward(x.numpy(), tgt.numpy(), eps.numpy(), 0.6)
        model.backward()
        d = _distance(out["x"], out["mu"], out["sigma"], out["recon"], out["kl"], out["loss"], model.gradients(), fw, grads, sp, T)
        report[operand] = d
        print("full-size parity [%s]: x_t err per step %s loss %.2e | grad rel-err median %.2e p95 %.2e max %.2e (%s) cosine "
              "median %.6f min %.6f   [fp32 CPU oracle: x_t %s grad median %.2e p95 %.2e]"
              % (operand, ["%.1e" % v for v in d["forward_max_rel_err_per_step"]["x"]], d["forward_max_rel_err_per_step"]["loss"],
                 d["grad_rel_err"]["median"], d["grad_rel_err"]["p95"], d["grad_rel_err"]["max"], d["grad_rel_err"]["worst_tensor"],
                 d["grad_cosine"]["median"], d["grad_cosine"]["min"],
                 ["%.1e" % v for v in cpu32["forward_max_rel_err_per_step"]["x"]], cpu32["grad_rel_err"]["median"],
                 cpu32["grad_rel_err"]["p95"]))
        if operand == "fp32":
            f, c = d["forward_max_rel_err_per_step"], cpu32["forward_max_rel_err_per_step"]
            for key in ("mu", "sigma", "x", "recon", "kl"):
                for t in range(T):
                    assert f[key][t] < max(1e-3, 2 * c[key][t]), (key, t, f[key][t], c[key][t])
            assert f["loss"] < max(1e-3, 2 * c["loss"])
            for q in ("median", "p95", "max"):
                assert d["grad_rel_err"][q] < max(1e-3, 2 * cpu32["grad_rel_err"][q]), (q, d["grad_rel_err"], cpu32["grad_rel_err"])
            assert d["grad_cosine"]["min"] > 1 - 2 * (1 - cpu32["grad_cosine"]["min"]) - 1e-6
        else:
            # What do bf16 operands cost on THIS graph, independent of any kernel?  The fp64 oracle with the operands of the same
            # contractions rounded to bf16 (fp64 accumulation): the chain's chaos turns the 2^-9 operand rounding into O(1)
            # differences of individual pixels and gradient entries by step 8 (the loss and the ELBO terms still agree to
            # 1e-4).  The device must sit at the same distance from fp64 as that emulation - no closer is possible, and
            # further would mean a kernel problem (which the local replay would show at 1e-4).
            from gpu_util import oracle_mode

            with oracle_mode("bf16"):
                fe, ge = O.loss_and_grads(hp, P, x, tgt, eps, 0.6)
            emu = _distance(fe["x"], fe["mu"], fe["sigma"], fe["recon"], fe["kl"], fe["loss"],
                            {k: v.numpy() for k, v in ge.items() if v is not None}, fw, grads, sp, T)
            report["bf16_operand_emulation_cpu_oracle"] = emu
            print("full-size parity [bf16 emulation on the CPU oracle]: x_t %s loss %.2e grad median %.2e cosine median %.4f"
                  % (["%.1e" % v for v in emu["forward_max_rel_err_per_step"]["x"]], emu["forward_max_rel_err_per_step"]["loss"],
                     emu["grad_rel_err"]["median"], emu["grad_cosine"]["median"]))
            f, c = d["forward_max_rel_err_per_step"], emu["forward_max_rel_err_per_step"]
            for key in ("mu", "sigma", "x", "recon", "kl"):
                for t in range(T):
                    assert f[key][t] < max(1e-3, 3 * c[key][t]), (key, t, f[key][t], c[key][t])
            assert f["loss"] < max(1e-3, 3 * c["loss"])
            assert d["grad_rel_err"]["median"] < max(1e-3, 2 * emu["grad_rel_err"]["median"])
            assert d["grad_rel_err"]["p95"] < max(1e-3, 2 * emu["grad_rel_err"]["p95"])
        model.close()
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "parity_fullsize.json"), "w") as fjson:
        json.dump(report, fjson, indent=1)


def test_full_size_local_replay_bf16():
    """Every tensor-core block of the production (bf16, TMA-fed) plan at the benchmarked size, replayed against the oracle
    layer on the device's own tensors (tests/replay.py): first, second and last chain step."""
    from gpu_util import make_inputs, make_pair
    from replay import replay
    from test_gpu_replay import check

    B = 100
    model, hp, P = make_pair("c_inhomog", [64, 64, 3], (-1.0, 1.0), B, "bf16")
    x, eps = make_inputs(hp, B)
    model.forward(x.numpy(), None, eps.numpy(), 0.6)
    model.backward()
    worst = replay(model, hp, P, "bf16", steps=[0, 1, hp["mc_steps"] - 1])
    print("full-size local replay bf16: %s" % {k: "%.2e @ %s" % v for k, v in worst.items()})
    check(worst, "bf16")
    model.close()
