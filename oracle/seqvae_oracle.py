"""CPU oracle for the Sequential-VAE hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain PyTorch-CPU restatement of the reference graph (MWPainter/Sequential-Variational-Autoencoder,
TensorFlow 1.x).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the *checker* or the *CPU baseline* - never as the product path.  The
product (``sequential-variational-autoencoder_b200``) must not import anything from ``oracle/``.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, it cannot be
imported here (TensorFlow 1.x / tf.contrib is not installable on Python 3.12), and the arithmetic lives
in a third-party dependency that is absent from /root/reference (TensorFlow 1.x, version unpinned; constrained to
1.4 <= TF < 2.0 by ``tf.AUTO_REUSE`` and ``tf.contrib``).  The TF-default semantics that are restated here (SAME
padding, conv2d_transpose == adjoint of conv, batch_norm defaults, xavier, TF-Adam) are listed in SURVEY.md App. B.
What *is* pinned: the SAME-padding/adjoint semantics are cross-checked against an independent naive numpy
implementation (``oracle/tf_semantics_np.py``), gradients against finite differences, and parameter counts
against the figures derived from the reference graph (782 tensors / 83 768 671 elements for CelebA).

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------
# Hyper-parameters (sequential_vae.py:201-258 defaults; netname overrides :281-862)
# --------------------------------------------------------------------------------------------------------------

_NETNAMES = {
    # netname -> overrides.  Only the rows of the netname table that are on the benchmarked path.
    "c_inhomog": {},                                    # sequential_vae.py:727
    "sequential_vae_celebA_inhomog": {},                # sequential_vae.py:671
    "sequential_vae_lsun": {"vlae_latent_dims": [20, 30, 30, 30]},   # sequential_vae.py:712-714
    "m_inhomog": {                                      # sequential_vae.py:842-848
        "vlae_levels": 3, "vlae_latent_dims": [2, 2, 2], "image_sizes": [32, 16, 8, 4],
        "filter_sizes": [None, 64, 128, 192, 256], "mc_steps": 5,
    },
    # homogeneous (weight-shared) chains and the other rows that only move knobs of the same path
    "sequential_vae_celebA_homog": {"share_theta_weights": True, "share_phi_weights": True},          # :675-677
    "sequential_vae_celebA_homog_fixed_length": {"share_theta_weights": True},                        # :709-710
    "c_homog": {"share_theta_weights": True, "share_phi_weights": True, "mc_steps": 25},              # :730-733
    "c_homog_v1": {"vlae_latent_dims": [12, 12, 12, 12], "filter_sizes": [None, 16, 32, 64, 128, 384],  # :316-321
                   "share_theta_weights": True, "share_phi_weights": True},
    "c_homog_one_step": {"vlae_latent_dims": [12, 12, 12, 12], "filter_sizes": [None, 16, 32, 64, 128, 384],  # :281-288
                         "share_theta_weights": True, "share_phi_weights": True, "mc_steps": 1},
    "s_homog_one_step": {"vlae_latent_dims": [12, 12, 12, 12], "share_theta_weights": True,           # :301-307
                         "share_phi_weights": True, "mc_steps": 1},
    "vlae_celebA": {"mc_steps": 1, "vlae_latent_dims": [16, 16, 16, 16]},                             # :704-707
    "sequential_vae_lsun_final": {"vlae_latent_dims": [20, 30, 30, 30], "intermediate_reconstruction": False},  # :721-725
    # chain noise with fixed per-step stddevs (first slice)
    "c_sample_images": {"add_noise_to_chain": True},                                                  # :761-762
    "c_homog_sample_images": {"share_theta_weights": True, "share_phi_weights": True, "add_noise_to_chain": True},  # :764-767
}


def hyperparams(netname: str, data_dims: Sequence[int], data_range=(0.0, 1.0), **overrides) -> dict:
    """Default hyper-parameters (sequential_vae.py:201-258) + the netname row (:281-862) + explicit overrides."""
    if netname not in _NETNAMES:
        raise KeyError("Unknown network name %s" % netname)      # sequential_vae.py:860-862 (exit(-1) there)
    D, C = int(data_dims[0]), int(data_dims[-1])
    hp = dict(
        name=netname, data_dims=[int(d) for d in data_dims], range=[float(data_range[0]), float(data_range[1])],
        vlae_levels=4, vlae_latent_dims=[3, 3, 3, 3],                                  # :202-203
        image_sizes=[D, D // 2, D // 4, D // 8, D // 16],                              # :204-206
        filter_sizes=[C, 32, 64, 128, 384, 512],                                       # :207
        mc_steps=8, intermediate_reconstruction=True,                                  # :219,221
        first_step_loss_coeff=1.0, latent_mean_clip=float("inf"), latent_prior_stddev=1.0,   # :226,229,231
        max_highway_ratio=1.0, min_highway_ratio=0.0,                                  # :240-241
        learning_rate=0.0002, learning_rate_decay=1.0, reg_coeff_rate=5000.0,          # :250-252
        clip_grads=True, clip_grad_value=10.0,                                         # :257-258
        share_theta_weights=False, share_phi_weights=False,                            # :213-214
        add_noise_to_chain=False,                                                      # :233
        noise_stddevs=[0.5 ** 1, 0.5 ** 2, 0.5 ** 3, 0.5 ** 4, 0.5 ** 5, 0.5 ** 6, 0.5 ** 7, 0],   # :239
    )
    hp["regularized_steps"] = list(range(hp["mc_steps"]))     # :224 - runs BEFORE the netname rows, i.e. with mc_steps == 8:
    row = dict(_NETNAMES[netname])                             # a row that lengthens the chain (c_homog, :733) keeps KL on steps 0..7
    if "filter_sizes" in row:
        row["filter_sizes"] = [C if f is None else f for f in row["filter_sizes"]]
    hp.update(row)
    hp.update(overrides)
    hp["latent_dim"] = int(np.sum(hp["vlae_latent_dims"]))                             # :220
    hp["regularized_steps"] = [int(t) for t in hp["regularized_steps"] if 0 <= int(t) < hp["mc_steps"]]   # `step in ...`, :1154
    L = hp["vlae_levels"]
    assert len(hp["image_sizes"]) == L + 1 and len(hp["filter_sizes"]) == L + 2 and len(hp["vlae_latent_dims"]) == L
    assert not hp["add_noise_to_chain"] or len(hp["noise_stddevs"]) == hp["mc_steps"]   # :151, indexed per step at :1736
    return hp


# --------------------------------------------------------------------------------------------------------------
# TF variable scoping: default layer names uniquify per scope entry (SURVEY App. B "Variable-scope naming")
# --------------------------------------------------------------------------------------------------------------

class Scope:
    """Hands out TF-contrib default layer names (Conv, Conv_1, BatchNorm, fully_connected, Conv2d_transpose ...)
    in call order inside one ``tf.variable_scope``; in spec mode it records (name, shape, init, flags) instead of
    fetching tensors."""

    def __init__(self, prefix: str, params: Optional[Dict[str, torch.Tensor]], specs: Optional[list]):
        self.prefix, self.params, self.specs = prefix, params, specs
        self.counts: Dict[str, int] = {}

    def layer(self, kind: str) -> str:
        n = self.counts.get(kind, 0)
        self.counts[kind] = n + 1
        return "%s/%s" % (self.prefix, kind if n == 0 else "%s_%d" % (kind, n))

    def get(self, name: str, shape, init: str, inert=False, dead=False, fan=None):
        if self.specs is not None:
            self.specs.append(dict(name=name, shape=tuple(int(s) for s in shape), init=init, inert=inert, dead=dead,
                                   fan=fan))
            return None
        return self.params[name]


# --------------------------------------------------------------------------------------------------------------
# Layer blocks (abstract_network.py:8-71) with the TF-contrib defaults of SURVEY App. B
# --------------------------------------------------------------------------------------------------------------

PREACT_HOOK = None  # tests may set a callable(tensor): called with every pre-activation (input of lrelu / relu)


def _probe(x):
    if PREACT_HOOK is not None:
        PREACT_HOOK(x.detach())
    return x


def lrelu(x, rate=0.1):
    """abstract_network.py:8-10: max(min(rate*x, 0), x)."""
    _probe(x)
    return torch.maximum(torch.clamp(x * rate, max=0.0), x)


def same_pads(size: int, k: int, s: int):
    """TF 'SAME': out=ceil(in/s); pad_total=max((out-1)s+k-in,0); before=total//2, rest after (App. B)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def _conv2d_same_exact(x, w_hwio, stride: int):
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    pt, pb = same_pads(x.shape[1], kh, stride)
    pl, pr = same_pads(x.shape[2], kw, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, w_hwio.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1)


def _conv2d_transpose_same_exact(x, w_hwoi, stride: int):
    kh, kw = w_hwoi.shape[0], w_hwoi.shape[1]
    H, W = x.shape[1], x.shape[2]
    pt, _ = same_pads(H * stride, kh, stride)
    pl, _ = same_pads(W * stride, kw, stride)
    w = w_hwoi.permute(3, 2, 0, 1)                       # [Cin_t, Cout_t, kh, kw] == adjoint of conv2d weight [O,I,kh,kw]
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w, stride=stride, padding=0)
    y = y[:, :, pt:pt + H * stride, pl:pl + W * stride]  # crop the forward conv's (before) padding
    return y.permute(0, 2, 3, 1)


# ---- operand-rounding emulation (verification of the bf16 tensor-core kernels only) -----------------------------------
# The CUDA library's SVAE_OPERAND_BF16 mode rounds the operands of the tensor-core contractions to bf16 (fp32
# accumulation).  OPERAND_EMULATION = dict(fwd=pred, dgrad=pred, wgrad=pred) makes the oracle round exactly those
# operands (pred(kind, h, w, cin, cout, stride) -> bool says whether the CUDA path runs that contraction on tensor cores) while
# keeping its own accumulation precision, so that the kernels can be checked tightly inside the full chain.  None (the
# default) is the reference semantics.
OPERAND_EMULATION = None


def _bf16(t):
    return t.to(torch.bfloat16).to(t.dtype)


class _EmulatedContraction(torch.autograd.Function):
    @staticmethod
    def _fn(transpose):
        if transpose == "fc":
            return lambda x, w, stride: x @ w
        return _conv2d_transpose_same_exact if transpose else _conv2d_same_exact

    @staticmethod
    def forward(ctx, x, w, stride, transpose, rf, rd, rw):
        fn = _EmulatedContraction._fn(transpose)
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, transpose, rd, rw)
        return fn(_bf16(x) if rf else x, _bf16(w) if rf else w, stride)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, transpose, rd, rw = ctx.cfg
        fn = _EmulatedContraction._fn(transpose)
        with torch.enable_grad():
            xd = x.detach().requires_grad_(True)
            dx, = torch.autograd.grad(fn(xd, (_bf16(w) if rd else w).detach(), stride), xd, _bf16(dy) if rd else dy)
            wd = w.detach().requires_grad_(True)
            dw, = torch.autograd.grad(fn((_bf16(x) if rw else x).detach(), wd, stride), wd, _bf16(dy) if rw else dy)
        return dx, dw, None, None, None, None, None


def _contract(x, w, stride, transpose):
    fn = _conv2d_transpose_same_exact if transpose else _conv2d_same_exact
    em = OPERAND_EMULATION
    if em is None:
        return fn(x, w, stride)
    cin = x.shape[-1]
    cout = w.shape[2] if transpose else w.shape[3]
    kind = "deconv" if transpose else "conv"
    rf, rd, rw = (bool(em[k](kind, int(x.shape[1]), int(x.shape[2]), int(cin), int(cout), stride))
                  for k in ("fwd", "dgrad", "wgrad"))
    if not (rf or rd or rw):
        return fn(x, w, stride)
    return _EmulatedContraction.apply(x, w, stride, transpose, rf, rd, rw)


def _fc(x, w):
    """x @ w (tf.contrib.layers.fully_connected data path), with the operand-rounding emulation when enabled."""
    em = OPERAND_EMULATION
    if em is None:
        return x @ w
    rf, rd, rw = (bool(em[k]("fc", 1, 1, int(w.shape[0]), int(w.shape[1]), 1)) for k in ("fwd", "dgrad", "wgrad"))
    if not (rf or rd or rw):
        return x @ w
    return _EmulatedContraction.apply(x, w, 1, "fc", rf, rd, rw)


def conv2d_same(x, w_hwio, stride: int):
    """tf.contrib.layers.convolution2d data path (abstract_network.py:18): NHWC, HWIO weights, SAME padding."""
    return _contract(x, w_hwio, stride, False)


def conv2d_transpose_same(x, w_hwoi, stride: int):
    """tf.contrib.layers.convolution2d_transpose data path (abstract_network.py:37,56; sequential_vae.py:1720,1727):
    weights [kh,kw,Cout,Cin]; computed by TF as conv2d_backprop_input, i.e. the exact adjoint of the SAME conv that
    maps [N,H*s,W*s,Cout] -> [N,H,W,Cin]."""
    return _contract(x, w_hwoi, stride, True)


def batch_norm(x, beta, eps=1e-3):
    """tf.contrib.layers.batch_norm(x) with ALL defaults (abstract_network.py:22 etc.): is_training=True always,
    scale=False (no gamma), center=True, epsilon=1e-3, biased batch variance; moving averages never used (Q1)."""
    dims = tuple(range(x.dim() - 1))
    mean = x.mean(dim=dims, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=dims, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) + beta


def _conv_block(sc: Scope, x, cin, cout, stride, act, transpose=False, dead=False):
    """conv2d_bn_lrelu (abstract_network.py:17-24) / conv2d_t_bn (:55-61) / conv2d_t_bn_relu (:36-43)."""
    lname = sc.layer("Conv2d_transpose" if transpose else "Conv")
    bname = sc.layer("BatchNorm")
    wshape = (4, 4, cout, cin) if transpose else (4, 4, cin, cout)
    w = sc.get(lname + "/weights", wshape, "normal0.02", dead=dead)
    b = sc.get(lname + "/biases", (cout,), "zeros", inert=True, dead=dead)
    beta = sc.get(bname + "/beta", (cout,), "zeros", dead=dead)
    if x is None:
        return None
    y = (conv2d_transpose_same(x, w, stride) if transpose else conv2d_same(x, w, stride)) + b
    y = batch_norm(y, beta)
    if act == "lrelu":
        y = lrelu(y)
    elif act == "relu":
        y = torch.relu(_probe(y))
    return y


def _fc_bn_lrelu(sc: Scope, x, nin, nout, dead=False):
    """fc_bn_lrelu (abstract_network.py:64-71)."""
    lname = sc.layer("fully_connected")
    bname = sc.layer("BatchNorm")
    w = sc.get(lname + "/weights", (nin, nout), "normal0.02", dead=dead)
    b = sc.get(lname + "/biases", (nout,), "zeros", inert=True, dead=dead)
    beta = sc.get(bname + "/beta", (nout,), "zeros", dead=dead)
    if x is None:
        return None
    return lrelu(batch_norm(_fc(x, w) + b, beta))


def _fc_head(sc: Scope, x, nin, nout):
    """layers.fully_connected(ladder, n, activation_fn=...) with default xavier init and a live bias
    (sequential_vae.py:1592,1594,1607,1609)."""
    lname = sc.layer("fully_connected")
    w = sc.get(lname + "/weights", (nin, nout), "xavier", fan=(nin, nout))
    b = sc.get(lname + "/biases", (nout,), "zeros")
    if x is None:
        return None
    return x @ w + b


def _deconv_out(sc: Scope, x, cin, cout, stride):
    """conv2d_t(cur_sample, n, [4,4], 2, activation_fn=tf.sigmoid): xavier init, live bias (sequential_vae.py:1720,1727)."""
    lname = sc.layer("Conv2d_transpose")
    w = sc.get(lname + "/weights", (4, 4, cout, cin), "xavier", fan=(16 * cout, 16 * cin))
    b = sc.get(lname + "/biases", (cout,), "zeros")
    if x is None:
        return None
    return torch.sigmoid(conv2d_transpose_same(x, w, stride) + b)


# --------------------------------------------------------------------------------------------------------------
# Networks (sequential_vae.py:1537-1808)
# --------------------------------------------------------------------------------------------------------------

def inference_ladder(hp, sc: Scope, x):
    """Recognition net x -> (mu_t, sigma_t), sequential_vae.py:1537-1630, INCLUDING the stale-`ladder` behaviour:
    the last level's conv + fc_bn_lrelu (:1602-1605) are dead code and the last heads read the level L-2 features
    (:1607,1609).  Dead variables are declared (spec mode) but never executed."""
    L, Fs, S, lat = hp["vlae_levels"], hp["filter_sizes"], hp["image_sizes"], hp["vlae_latent_dims"]
    clip = hp["latent_mean_clip"]
    cur, ladder = x, None
    mus, sds = [], []
    for level in range(L - 1):                                                       # :1585
        hidden = _conv_block(sc, cur, Fs[level], Fs[level + 1], 2, "lrelu")          # :1587
        cur = _conv_block(sc, hidden, Fs[level + 1], Fs[level + 1], 1, "lrelu")      # :1588
        nflat = S[level + 1] * S[level + 1] * Fs[level + 1]
        ladder = cur.reshape(cur.shape[0], -1) if x is not None else None            # :1591 (H,W,C order)
        mu = _fc_head(sc, ladder, nflat, lat[level])                                 # :1592
        sd = _fc_head(sc, ladder, nflat, lat[level])                                 # :1594
        if x is not None:
            mus.append(torch.clamp(mu, -clip, clip))                                 # :1593
            sds.append(torch.sigmoid(sd))
    # dead branch (:1602-1605): variables exist, result unused
    _conv_block(sc, None, Fs[L - 1], Fs[L - 1], 2, "lrelu", dead=True)
    _fc_bn_lrelu(sc, None, S[L] * S[L] * Fs[L - 1], Fs[L], dead=True)
    nflat = S[L - 1] * S[L - 1] * Fs[L - 1]
    mu = _fc_head(sc, ladder, nflat, lat[L - 1])                                     # :1607 (reads stale `ladder`)
    sd = _fc_head(sc, ladder, nflat, lat[L - 1])                                     # :1609
    if x is None:
        return None, None
    mus.append(torch.clamp(mu, -clip, clip))                                         # :1608
    sds.append(torch.sigmoid(sd))
    return torch.cat(mus, 1), torch.cat(sds, 1)                                      # :1630


def compute_encodings(hp, sc: Scope, x):
    """Chain encoder x_{t-1} -> [e0..eL], sequential_vae.py:1745-1777."""
    L, Fs, S = hp["vlae_levels"], hp["filter_sizes"], hp["image_sizes"]
    run = x is not None
    cur = x
    enc = [cur]
    for level in range(L - 1):                                                       # :1767
        hidden = _conv_block(sc, cur, Fs[level], Fs[level + 1], 2, "lrelu")          # :1768
        cur = _conv_block(sc, hidden, Fs[level + 1], Fs[level + 1], 1, "lrelu")      # :1769
        enc.append(cur)
    cur = _conv_block(sc, cur, Fs[L - 1], Fs[L - 1], 2, "lrelu")                     # :1772
    flat = cur.reshape(cur.shape[0], -1) if run else None                            # :1773
    cur = _fc_bn_lrelu(sc, flat, S[L] * S[L] * Fs[L - 1], Fs[L])                     # :1774
    enc.append(cur)
    return enc if run else None


def generator_ladder(hp, sc: Scope, enc, z, first_step: bool, declare_only=False):
    """Decoder (x_{t-1} encodings | None, z_t) -> x_t, sequential_vae.py:1636-1739 with split_latent (:1783-1808) and
    combine_noise 'concat' (:1833-1834, order [feature, plane])."""
    L, Fs, S, lat = hp["vlae_levels"], hp["filter_sizes"], hp["image_sizes"], hp["vlae_latent_dims"]
    C = hp["data_dims"][-1]
    lo, hi = hp["range"]
    run = not declare_only
    # split_latent (:1796-1806)
    groups = list(torch.split(z, lat, dim=1)) if run else [None] * L
    planes = []
    for i in range(L - 1):
        n = S[i + 1] * S[i + 1] * Fs[i + 1]
        p = _fc_bn_lrelu(sc, groups[i], lat[i], n)                                   # :1803
        planes.append(p.reshape(-1, S[i + 1], S[i + 1], Fs[i + 1]) if run else None)  # :1804
    planes.append(_fc_bn_lrelu(sc, groups[L - 1], lat[L - 1], Fs[L + 1]))            # :1805
    # first decoder level (:1695-1705)
    if not first_step:
        cur = torch.cat([enc[L], planes[L - 1]], 1) if run else None                 # :1696-1697
        nin = Fs[L] + Fs[L + 1]
    else:
        cur = planes[L - 1]                                                          # :1699
        nin = Fs[L + 1]
    cur = _fc_bn_lrelu(sc, cur, nin, S[L] * S[L] * Fs[L])                            # :1704
    cur = cur.reshape(-1, S[L], S[L], Fs[L]) if run else None                        # :1705
    cin = Fs[L]
    for level in range(L - 2, -1, -1):                                               # :1710
        d = _conv_block(sc, cur, cin, Fs[level + 1], 2, None, transpose=True)        # :1711
        if run:
            if not first_step:
                d = d + enc[level + 1]                                               # :1713
            d = torch.relu(_probe(d))                                                # :1714
            d = torch.cat([d, planes[level]], 3)                                     # :1716
        cur = _conv_block(sc, d if run else None, 2 * Fs[level + 1], Fs[level + 1], 1, "relu", transpose=True)  # :1717
        cin = Fs[level + 1]
    o = _deconv_out(sc, cur, cin, C, 2)                                              # :1720
    out = (hi - lo) * o + lo if run else None                                        # :1721
    ratio = None
    if not first_step:
        r = _deconv_out(sc, cur, cin, 1, 2)                                          # :1727
        if run:
            ratio = hp["min_highway_ratio"] + (hp["max_highway_ratio"] - hp["min_highway_ratio"]) * r.expand(-1, -1, -1, C)
            out = ratio * out + (1 - ratio) * enc[0]                                 # :1728-1729
    return out, ratio


# --------------------------------------------------------------------------------------------------------------
# Parameter table (TF trainable_variables() creation order)
# --------------------------------------------------------------------------------------------------------------

def phi_scope(hp, t: int) -> str:
    """sequential_vae.py:1573-1577: one scope per step, or "phi/inference_network" for every step when phi is shared
    (tf.AUTO_REUSE; the default layer names restart on every scope entry, so the same variables are found again)."""
    return "phi/inference_network" if hp.get("share_phi_weights") else "phi/inference_step_%d" % t


def encoder_scope(hp, t: int) -> str:
    """sequential_vae.py:1757-1761."""
    return "theta/generative_encoder_network" if hp.get("share_theta_weights") else "theta/generative_encoder_step_%d" % t


def generator_scope(hp, t: int) -> str:
    """sequential_vae.py:1683-1687: shared only when the step has a chain input; step 0 keeps its own decoder."""
    return "theta/generative_network" if hp.get("share_theta_weights") and t > 0 else "theta/generative_step_%d" % t


def param_specs(hp) -> List[dict]:
    """All trainable variables in TF creation order: per step, phi/inference_step_t, then (t>=1)
    theta/generative_encoder_step_t, then theta/generative_step_t (construct_network, sequential_vae.py:934-975)."""
    specs: List[dict] = []
    for t in range(hp["mc_steps"]):
        inference_ladder(hp, Scope(phi_scope(hp, t), None, specs), None)
        if t > 0:
            compute_encodings(hp, Scope(encoder_scope(hp, t), None, specs), None)
        generator_ladder(hp, Scope(generator_scope(hp, t), None, specs), None, None, t == 0, declare_only=True)
    seen, uniq = set(), []
    for sp in specs:            # a shared scope re-entered with tf.AUTO_REUSE finds its variables, it does not create new ones
        if sp["name"] not in seen:
            seen.add(sp["name"])
            uniq.append(sp)
    return uniq


def init_params(hp, seed=0, dtype=torch.float64) -> "OrderedDict[str, torch.Tensor]":
    """Reference initialisers (App. B): N(0,0.02) for *_bn_* blocks (abstract_network.py:19 etc.), xavier-uniform for
    heads / output deconvs (contrib default), zeros for biases and betas.  Drawn in fp32 then cast so that the fp64 and
    fp32 oracles (and the CUDA path, which is fed these same arrays) share identical weights."""
    g = torch.Generator().manual_seed(seed)
    P: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for s in param_specs(hp):
        if s["init"] == "zeros":
            v = torch.zeros(s["shape"], dtype=torch.float32)
        elif s["init"] == "normal0.02":
            v = torch.randn(s["shape"], generator=g, dtype=torch.float32) * 0.02
        elif s["init"] == "xavier":
            lim = math.sqrt(6.0 / (s["fan"][0] + s["fan"][1]))
            v = (torch.rand(s["shape"], generator=g, dtype=torch.float32) * 2 - 1) * lim
        else:
            raise ValueError(s["init"])
        P[s["name"]] = v.to(dtype)
    return P


# --------------------------------------------------------------------------------------------------------------
# Chain, loss, optimiser (sequential_vae.py:877-1093, 1101-1212, 1246-1320)
# --------------------------------------------------------------------------------------------------------------

def forward_chain(hp, P, x_in, x_target, eps, reg_coeff=1.0, chain_eps=None):
    """Training-mode chain (construct_network, sequential_vae.py:934-975): for every step a recognition net on x_in
    (:1022), z = mu + sigma*eps (:1023, eps injected), the generator on (x_{t-1}, z) with gradient flowing through
    x_{t-1} (Q4: the stop_gradient at :1211-1212 is a no-op), and the per-step ELBO terms (:1146-1176).
    eps: [T,B,Z].  add_noise_to_chain (:1088-1091): the SAMPLE x_t + reg_coeff * noise_stddevs[t] * chain_eps[t]
    (chain_eps: [T,B,H,W,C], the injected tf.random_normal) is what step t+1 reads (:936,958); the losses stay on the mle.
    Returns a dict of per-step tensors and the total loss."""
    T = hp["mc_steps"]
    prior = hp["latent_prior_stddev"]
    noisy = bool(hp.get("add_noise_to_chain", False))
    out = dict(mu=[], sigma=[], z=[], x=[], ratio=[], recon=[], kl=[], sample=[])
    loss = 0.0
    prev = None
    for t in range(T):
        mu, sd = inference_ladder(hp, Scope(phi_scope(hp, t), P, None), x_in)
        z = mu + sd * eps[t]                                                         # :1023
        enc = compute_encodings(hp, Scope(encoder_scope(hp, t), P, None), prev) if t > 0 else None
        xt, ratio = generator_ladder(hp, Scope(generator_scope(hp, t), P, None), enc, z, t == 0)
        recon = ((xt - x_target) ** 2).mean(dim=(1, 2, 3)).mean()                    # :1146,1163
        kl = (-0.5 - torch.log(sd) + 0.5 * sd ** 2 / prior ** 2 + 0.5 * mu ** 2 / prior ** 2).mean(dim=1).mean()  # :1156-1164
        if hp["intermediate_reconstruction"] or t == T - 1:
            loss = loss + 16 * recon                                                 # :1167-1168
        if t in hp["regularized_steps"]:
            loss = loss + reg_coeff * kl                                             # :1171-1172
        if t == 0:
            loss = loss * hp["first_step_loss_coeff"]                                # :1175-1176
        sample = xt                                                                  # training_sample == mle (:1090, stddev 0)
        if noisy:
            sample = xt + reg_coeff * hp["noise_stddevs"][t] * chain_eps[t]          # :1090 (stddevs = noise_stddevs[step], :1736)
        for k, v in (("mu", mu), ("sigma", sd), ("z", z), ("x", xt), ("ratio", ratio), ("recon", recon), ("kl", kl),
                     ("sample", sample)):
            out[k].append(v)
        prev = sample                                                                # prev_training_sample (:936)
    out["loss"] = loss
    out["final_loss"] = out["recon"][-1]                                             # :1204
    return out


def generate_chain(hp, P, z, batch_size, chain_eps=None, samples_out=None):
    """Generation-mode chain (generative twins, sequential_vae.py:1068-1073,1397-1428): z[t] fed from the host;
    recognition nets not evaluated; BN uses the generated batch's statistics (Q1).  Returns the mles [x_1..x_T] (the
    reference additionally prepends a uniform-noise x_0 that nothing depends on, :947-952).  add_noise_to_chain: step t+1 reads
    generative_sample = mle + reg_coeff * noise_stddevs[t] * chain_eps[t] with the placeholder's default reg_coeff = 1 (:917,1091);
    the samples are appended to `samples_out`."""
    xs, prev = [], None
    noisy = bool(hp.get("add_noise_to_chain", False))
    for t in range(hp["mc_steps"]):
        enc = compute_encodings(hp, Scope(encoder_scope(hp, t), P, None), prev) if t > 0 else None
        xt, _ = generator_ladder(hp, Scope(generator_scope(hp, t), P, None), enc, z[t], t == 0)
        xs.append(xt)
        prev = xt + hp["noise_stddevs"][t] * chain_eps[t] if noisy else xt
        if samples_out is not None:
            samples_out.append(prev)
    return xs


def loss_and_grads(hp, P, x_in, x_target, eps, reg_coeff=1.0, chain_eps=None):
    """Forward + reverse-mode through the whole chain (optimizer.compute_gradients, sequential_vae.py:1273).
    Returns (forward dict, {name: grad or None}); dead-branch variables get None like in TF."""
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in P.items())
    fw = forward_chain(hp, leaves, x_in, x_target, eps, reg_coeff, chain_eps)
    names = list(leaves.keys())
    gs = torch.autograd.grad(fw["loss"], [leaves[n] for n in names], allow_unused=True)
    grads = OrderedDict(zip(names, gs))
    fw = {k: ([None if u is None else u.detach() for u in v] if isinstance(v, list) else v.detach()) for k, v in fw.items()}
    return fw, grads


class AdamState:
    """tf.train.AdamOptimizer slots (m, v per variable; beta powers), SURVEY App. B / Q11."""

    def __init__(self, P):
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
        self.t = 0


def adam_apply(hp, P, grads, st: AdamState, lr, beta1=0.9, beta2=0.999, eps=1e-8, update_inert=True):
    """clip_by_value(+-clip_grad_value) (sequential_vae.py:18-25,1275) then TF-formulation Adam (:1267,1276):
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps).  Variables with None grads are skipped."""
    st.t += 1
    lr_t = lr * math.sqrt(1 - beta2 ** st.t) / (1 - beta1 ** st.t)
    inert = {s["name"] for s in param_specs(hp) if s["inert"]}
    for k, g in grads.items():
        if g is None or (not update_inert and k in inert):
            continue
        if hp["clip_grads"]:
            g = torch.clamp(g, -hp["clip_grad_value"], hp["clip_grad_value"])
        st.m[k].mul_(beta1).add_(g, alpha=1 - beta1)
        st.v[k].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        P[k].sub_(lr_t * st.m[k] / (st.v[k].sqrt() + eps))


class OracleModel:
    """Mirror of the SequentialVAE run wrappers (sequential_vae.py:1341-1455) on the oracle graph."""

    def __init__(self, hp, seed=0, dtype=torch.float64):
        self.hp, self.dtype = hp, dtype
        self.P = init_params(hp, seed, dtype)
        self.adam = AdamState(self.P)
        self.iteration = 0
        self.learning_rate = hp["learning_rate"]

    def train(self, input_batch, batch_target, eps, update_inert=True):
        """SequentialVAE.train (sequential_vae.py:1341-1375): schedules, one fwd+bwd+Adam, returns final_loss/H/W."""
        hp = self.hp
        self.iteration += 1
        self.learning_rate *= hp["learning_rate_decay"]
        reg = 1 - math.exp(-self.iteration / hp["reg_coeff_rate"])                    # :1357
        x = torch.as_tensor(input_batch, dtype=self.dtype)
        tgt = torch.as_tensor(batch_target, dtype=self.dtype)
        fw, grads = loss_and_grads(hp, self.P, x, tgt, torch.as_tensor(eps, dtype=self.dtype), reg)
        with torch.no_grad():
            adam_apply(hp, self.P, grads, self.adam, self.learning_rate, update_inert=update_inert)
        return float(fw["final_loss"]) / hp["data_dims"][0] / hp["data_dims"][1], fw, grads   # :1375

    def test(self, input_batch, eps):
        """SequentialVAE.test (sequential_vae.py:1381-1391): training-mode chain with reg_coeff default 1.0 -> x_T."""
        x = torch.as_tensor(input_batch, dtype=self.dtype)
        with torch.no_grad():
            return forward_chain(self.hp, self.P, x, x, torch.as_tensor(eps, dtype=self.dtype), 1.0)["x"][-1]

    def generate_mc_samples(self, z):
        """SequentialVAE.generate_mc_samples (sequential_vae.py:1397-1428) minus the leading uniform-noise x_0."""
        z = torch.as_tensor(z, dtype=self.dtype)
        with torch.no_grad():
            return generate_chain(self.hp, self.P, z, z.shape[1])


# --------------------------------------------------------------------------------------------------------------
# Host step in front of the path: denoising corruption (trainer.py:56-78)
# --------------------------------------------------------------------------------------------------------------

def apply_noise(original, keep, salt, gauss, data_range):
    """NoisyTrainer.apply_noise (trainer.py:56-78) with the three random fields passed in instead of drawn from numpy's
    unseeded global RNG: keep ~ binomial(1, 1 - pepper_prob) (:69), salt ~ binomial(1, salt_prob) (:70),
    gauss ~ normal(scale=gaussian_noise_scale) (:73); the result is clipped to dataset.range (:76)."""
    noisy = np.multiply(original, keep) + salt
    noisy = noisy + gauss
    return np.clip(noisy, a_min=data_range[0], a_max=data_range[1])


NOISE_DEFAULTS = dict(pepper_prob=0.1, salt_prob=0.1, gaussian_noise_scale=0.1)   # trainer.py:16-18
