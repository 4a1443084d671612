"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/svae.h declares, fails loudly without
a GPU, and its device-free parameter table equals the oracle's (TF creation order, shapes, inert/dead/xavier flags)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import seqvae_b200 as S
from seqvae_b200 import _cabi
from seqvae_b200.config import to_cabi_config
from oracle import seqvae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "svae.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _cabi.lib()
    syms = _declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(L, s), "libsvae.so does not export %s" % s
        assert s in _cabi.PROTOTYPES, "ctypes binding has no prototype for %s" % s
    assert set(_cabi.PROTOTYPES) == set(syms)
    assert b"sm_100a" in L.svae_version()


def test_struct_layouts_match_header():
    # sizes implied by include/svae.h (no padding surprises between C and ctypes)
    assert C.sizeof(_cabi.Losses) == 4 * (2 + 2 * 64)
    assert C.sizeof(_cabi.ParamInfo) == 128 + 4 + 16 + 4 + 8 + 8 + 4 + 4 or C.sizeof(_cabi.ParamInfo) % 8 == 0
    assert C.sizeof(_cabi.Config) % 8 == 0


@pytest.mark.parametrize("name,dims,rng", [("c_inhomog", [64, 64, 3], (-1, 1)), ("m_inhomog", [32, 32, 1], (0, 1)),
                                           ("sequential_vae_lsun", [64, 64, 3], (-1, 1)),
                                           ("sequential_vae_celebA_inhomog", [64, 64, 3], (-1, 1)),
                                           ("sequential_vae_celebA_homog", [64, 64, 3], (-1, 1)),
                                           ("sequential_vae_celebA_homog_fixed_length", [64, 64, 3], (-1, 1)),
                                           ("c_homog", [64, 64, 3], (-1, 1)), ("c_homog_v1", [32, 32, 3], (0, 1)),
                                           ("c_homog_one_step", [32, 32, 3], (0, 1)), ("vlae_celebA", [64, 64, 3], (-1, 1)),
                                           ("sequential_vae_lsun_final", [64, 64, 3], (-1, 1))])
def test_param_table_matches_oracle(name, dims, rng):
    L = _cabi.lib()
    cfg = to_cabi_config(S.hyperparams(name, dims, rng), 100)
    n = L.svae_param_table(C.byref(cfg), None, 0)
    arr = (_cabi.ParamInfo * n)()
    assert L.svae_param_table(C.byref(cfg), arr, n) == n
    sp = O.param_specs(O.hyperparams(name, dims, rng))
    assert n == len(sp)
    off = 0
    for a, s in zip(arr, sp):
        assert a.name.decode() == s["name"]
        assert tuple(a.shape[:a.ndim]) == s["shape"]
        assert a.numel == int(np.prod(s["shape"]))
        assert bool(a.flags & _cabi.PF_INERT) == s["inert"]
        assert bool(a.flags & _cabi.PF_DEAD) == s["dead"]
        assert bool(a.flags & _cabi.PF_XAVIER) == (s["init"] == "xavier")
        assert bool(a.flags & _cabi.PF_THETA) == s["name"].startswith("theta/")
        assert a.offset >= off and a.offset % 4 == 0
        off = a.offset + a.numel
        m = re.search(r"step_(\d+)", s["name"])
        if m:
            assert a.step == int(m.group(1))
        else:       # shared scope (homogeneous chain): owned by the first step that uses it
            assert a.step == (0 if s["name"].startswith("phi/inference_network") else 1)


def test_invalid_configs_are_rejected():
    L = _cabi.lib()
    hp = S.hyperparams("c_inhomog", [64, 64, 3], (-1, 1))
    cfg = to_cabi_config(hp, 100)
    cfg.height = 60                                   # not a multiple of 2^levels
    cfg.width = 60
    assert L.svae_param_table(C.byref(cfg), None, 0) == -1
    cfg = to_cabi_config(hp, 100)
    cfg.latent_dims[1] = 64                           # above the supported latent width
    assert L.svae_param_table(C.byref(cfg), None, 0) == -1
    cfg = to_cabi_config(hp, 0)
    assert L.svae_param_table(C.byref(cfg), None, 0) == -1


def test_unknown_netname_and_bad_image_sizes():
    with pytest.raises(KeyError):
        S.hyperparams("no_such_net", [64, 64, 3], (-1, 1))          # reference: log error + exit(-1)
    with pytest.raises(ValueError):
        S.hyperparams("m_inhomog", [64, 64, 1], (0, 1))             # image_sizes [32,16,8,4] vs 64x64 input (:1617-1627)


def test_shared_scopes_hold_one_copy():
    """sequential_vae.py:1573-1577,1683-1687,1757-1761: a homogeneous chain has ONE recognition net, ONE chain encoder and
    ONE decoder for steps >= 1 (+ step 0's own decoder), whatever the chain length."""
    names = lambda hp: [s["name"] for s in O.param_specs(hp)]
    a = names(O.hyperparams("sequential_vae_celebA_homog", [64, 64, 3], (-1, 1)))
    b = names(O.hyperparams("c_homog", [64, 64, 3], (-1, 1)))                      # same nets, 25 steps
    assert a == b and len(set(a)) == len(a)
    scopes = []
    for n in a:
        sc = "/".join(n.split("/")[:2])
        if sc not in scopes:
            scopes.append(sc)
    assert scopes == ["phi/inference_network", "theta/generative_step_0", "theta/generative_encoder_network",
                      "theta/generative_network"]
    inh = O.param_specs(O.hyperparams("c_inhomog", [64, 64, 3], (-1, 1)))
    per_step = [s for s in inh if "_step_1/" in s["name"]]
    step0_dec = [s for s in inh if s["name"].startswith("theta/generative_step_0/")]
    assert len(a) == len(per_step) + len(step0_dec)


def test_netname_rows():
    hp = S.hyperparams("m_inhomog", [32, 32, 1], (0, 1))
    assert (hp["vlae_levels"], hp["mc_steps"], hp["latent_dim"]) == (3, 5, 6)
    assert hp["filter_sizes"] == [1, 64, 128, 192, 256]
    hp = S.hyperparams("sequential_vae_lsun", [64, 64, 3], (-1, 1))
    assert hp["latent_dim"] == 110 and hp["mc_steps"] == 8
    hp = S.hyperparams("c_inhomog", [64, 64, 3], (-1, 1))
    assert hp["filter_sizes"] == [3, 32, 64, 128, 384, 512] and hp["learning_rate"] == 2e-4
    for name in S.NETNAMES:                                                      # product table == oracle table
        hp = S.hyperparams(name, [32, 32, 3] if name.startswith("m_") is False else [32, 32, 1], (-1, 1))
        for k, v in O.hyperparams(name, hp["data_dims"], (-1, 1)).items():
            assert hp[k] == v, (name, k)
    assert set(S.NETNAMES) == set(O._NETNAMES)
    hp = S.hyperparams("c_homog", [64, 64, 3], (-1, 1))
    assert hp["mc_steps"] == 25 and hp["share_theta_weights"] and hp["share_phi_weights"]
    # regularized_steps = range(mc_steps) is evaluated at sequential_vae.py:224, before the netname row lengthens the chain (:733)
    assert hp["regularized_steps"] == list(range(8))
    assert to_cabi_config(hp, 4).regularized_mask == 0xFF
    assert S.hyperparams("m_inhomog", [32, 32, 1], (0, 1))["regularized_steps"] == list(range(5))
    assert S.hyperparams("c_inhomog", [64, 64, 3], (-1, 1), mc_steps=3, regularized_steps=[0, 2, 7])["regularized_steps"] == [0, 2]
    assert not S.hyperparams("sequential_vae_celebA_homog_fixed_length", [64, 64, 3], (-1, 1))["share_phi_weights"]


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ds = S.SyntheticDataset("celebA", 4)
    with pytest.raises(_cabi.SvaeError) as ei:
        S.SequentialVAE(ds, 4, "c_inhomog")
    assert ei.value.code == -2


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sequential-variational-autoencoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_synthetic_dataset_shapes():
    ds = S.SyntheticDataset("mnist", 7)
    b = ds.next_batch(7)
    assert b.shape == (7, 32, 32, 1) and b.dtype == np.float32 and 0 <= b.min() and b.max() <= 1
    ds = S.SyntheticDataset("celebA", 3)
    b = ds.next_test_batch(3)
    assert b.shape == (3, 64, 64, 3) and -1 <= b.min() and b.max() <= 1
    ds.reset()
    np.testing.assert_array_equal(ds.next_test_batch(3), b)


def test_checkpoint_layout_is_the_tf_saver_variable_set():
    """abstract_network.py:124-152: tf.train.Saver() writes every global variable - trainables, BN moving statistics,
    Adam slots of the variables that receive a gradient (not the dead branch), beta powers."""
    from seqvae_b200.checkpoint import adam_t_from_beta_powers, beta_powers, tf_checkpoint_layout

    L = _cabi.lib()
    for name, shared in (("c_inhomog", False), ("sequential_vae_celebA_homog", True)):
        cfg = to_cabi_config(S.hyperparams(name, [64, 64, 3], (-1, 1)), 100)
        n = L.svae_param_table(C.byref(cfg), None, 0)
        arr = (_cabi.ParamInfo * n)()
        L.svae_param_table(C.byref(cfg), arr, n)
        table = [dict(name=a.name.decode(), shape=tuple(a.shape[:a.ndim]), flags=a.flags) for a in arr]
        lay = tf_checkpoint_layout(table, train=True)
        names = [e[0] for e in lay]
        assert len(set(names)) == len(names)
        kinds = {}
        for e in lay:
            kinds[e[2]] = kinds.get(e[2], 0) + 1
        n_bn = sum(1 for t in table if t["name"].endswith("/beta"))
        n_dead = sum(1 for t in table if t["flags"] & _cabi.PF_DEAD)
        assert kinds["param"] == n and kinds["bn_moving_mean"] == kinds["bn_moving_variance"] == n_bn
        assert kinds["adam_m"] == kinds["adam_v"] == n - n_dead and kinds["beta1_power"] == kinds["beta2_power"] == 1
        if not shared:
            assert n == 782 and n_dead == 6 * 8        # 782 trainable tensors; 6 dead-branch variables per step
            assert "phi/inference_step_3/BatchNorm_2/moving_variance" in names
            assert "theta/generative_step_7/Conv2d_transpose_7/weights/Adam_1" in names
        else:
            assert "theta/generative_network/BatchNorm/moving_mean" in names and not any("_step_1/" in x for x in names)
        assert len(tf_checkpoint_layout(table, train=False)) == n + 2 * n_bn
    for t in (0, 1, 7, 250):
        assert adam_t_from_beta_powers(*beta_powers(t)) == t
        assert adam_t_from_beta_powers(beta_powers(t)[0], None) == t
    # a real TF checkpoint: the fp32 beta1_power has underflowed to 0, beta2_power still resolves the step count
    for t in (1500, 20000, 80000):
        b1p, b2p = (np.float32(v) for v in beta_powers(t))
        assert b1p == 0.0 and abs(adam_t_from_beta_powers(b1p, b2p) - t) <= max(2, t * 2e-4)
    assert adam_t_from_beta_powers(np.float32(0.0), np.float32(0.0)) == 1000000


def test_noisy_trainer_host_logic_with_a_fake_network():
    """NoisyTrainer (trainer.py:82-141) is host logic only: which network calls it makes, when it tests / visualises / logs,
    what error it reports - checked here without a GPU against a recording stand-in for SequentialVAE."""
    import argparse
    import logging

    class FakeNet:
        def __init__(self):
            self.calls = []

        def train(self, x, tgt):
            self.calls.append(("train", x is tgt))
            return 0.5

        def train_denoise(self, tgt, pepper, salt, scale):
            self.calls.append(("train_denoise", (pepper, salt, scale)))
            return 0.25

        def apply_noise(self, x, pepper, salt, scale, seed=0):
            self.calls.append(("apply_noise", seed))
            return x + 1.0

        def test(self, x):
            self.calls.append(("test", float(x.mean())))
            return x                                   # reconstruction == (noisy) input

        def visualize(self, epoch):
            self.calls.append(("visualize", epoch))

    ds = S.SyntheticDataset("x", 4, data_dims=[8, 8, 3], data_range=[0.0, 1.0])
    for denoise in (False, True):
        net = FakeNet()
        args = argparse.Namespace(batch_size=4, denoise_train=denoise, vis_frequency=3, plot_reconstruction=False)
        tr = S.NoisyTrainer(net, ds, args, logging.getLogger("t"), "unused")
        loss = tr.train(max_iters=7)
        kinds = [c[0] for c in net.calls]
        assert kinds.count("visualize") == 3 and [c[1] for c in net.calls if c[0] == "visualize"] == [0, 1, 2]   # iterations 0, 3, 6
        assert kinds.count("test") == 3 * S.NoisyTrainer.test_num_iters
        if denoise:
            assert loss == 0.25 and kinds.count("train_denoise") == 7 and kinds.count("train") == 0
            assert ("train_denoise", (0.1, 0.1, 0.1)) in net.calls                     # trainer.py:16-18
            seeds = [c[1] for c in net.calls if c[0] == "apply_noise"]
            assert len(seeds) == 15 and len(set(seeds)) == 15                          # a fresh Philox key per corrupted test batch
            # reconstruction = input + 1 everywhere -> error per pixel = sum over C of 1 = 3 (trainer.py:131)
            assert abs(tr.test(0, num_iters=2) - 3.0) < 1e-5
        else:
            assert loss == 0.5 and kinds.count("train") == 7 and all(c[1] for c in net.calls if c[0] == "train")
            assert tr.test(0, num_iters=1) == 0.0
            with pytest.raises(Exception, match="denoise_train==False"):
                tr.apply_noise(ds.next_batch(4))                                       # trainer.py:66-67
