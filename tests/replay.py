"""Local replay of every tensor-core block of the REAL chain against the oracle layer (test infrastructure).

The whole-chain comparison with the fp64 oracle cannot give a tight bound for the bf16 operand family: rounding an
operand to bf16 moves a near-zero pre-activation across zero, and from there the two evaluations follow different ReLU
branches (chaos, not error).  The local replay removes the amplification: after one forward + backward of the production
plan, every block's own inputs are read back from the device (``SequentialVAE.block_tensor``) and the ORACLE LAYER
(abstract_network.py:17-71 restated in oracle/seqvae_oracle.py) is evaluated on exactly those inputs:

  forward   y   = contraction(x_gpu, W)                       vs the device's pre-batch-norm tensor
            out = act(batch_norm(y_gpu) + shortcut)           vs the device's activated output
  backward  dy  = d/dy [act(batch_norm(y) + shortcut)] . da_gpu   vs the device's dL/dy
            dbeta, dW = wgrad(x_gpu, dy_gpu)                  vs the gradient arena
            dx  = dgrad(dy_gpu, W) (+ the other contributions of that tensor)   vs the device's dL/d(out) of the producer

so every production kernel launch of the step (tc2_conv forward / input gradient, tc2_wgrad, the batch-norm kernels, the fc
kernels) is compared with the reference arithmetic on identical operands, inside the real chain.
"""
import numpy as np
import torch

from oracle import seqvae_oracle as O
from gpu_util import oracle_mode


def _nm(base, n):
    return base if n == 0 else "%s_%d" % (base, n)


def block_names(hp, t):
    """{(net, index): (weights name, beta name | None, kind, stride, act)} of chain step t, TF default layer names in call
    order inside each scope (oracle.Scope)."""
    L = hp["vlae_levels"]
    out = {}
    phi = O.phi_scope(hp, t)
    for k in range(2 * (L - 1)):
        out[("inf", k)] = (phi + "/" + _nm("Conv", k) + "/weights", phi + "/" + _nm("BatchNorm", k) + "/beta", "conv",
                           2 if k % 2 == 0 else 1, "lrelu")
    if t > 0:
        sc = O.encoder_scope(hp, t)
        for k in range(2 * (L - 1) + 1):
            out[("enc", k)] = (sc + "/" + _nm("Conv", k) + "/weights", sc + "/" + _nm("BatchNorm", k) + "/beta", "conv",
                               2 if k % 2 == 0 else 1, "lrelu")
        out[("encfc", 0)] = (sc + "/fully_connected/weights", sc + "/" + _nm("BatchNorm", 2 * (L - 1) + 1) + "/beta", "fc", 1,
                             "lrelu")
    sc = O.generator_scope(hp, t)
    out[("decfc", 0)] = (sc + "/" + _nm("fully_connected", L) + "/weights", sc + "/" + _nm("BatchNorm", L) + "/beta", "fc", 1,
                         "lrelu")
    for level in range(L - 2, -1, -1):
        j = 2 * (L - 2 - level)
        out[("ta", level)] = (sc + "/" + _nm("Conv2d_transpose", j) + "/weights",
                              sc + "/" + _nm("BatchNorm", L + 1 + j) + "/beta", "deconv", 2, "relu")
        out[("tb", level)] = (sc + "/" + _nm("Conv2d_transpose", j + 1) + "/weights",
                              sc + "/" + _nm("BatchNorm", L + 2 + j) + "/beta", "deconv", 1, "relu")
    out[("out", 0)] = (sc + "/" + _nm("Conv2d_transpose", 2 * (L - 1)) + "/weights", None, "deconv", 2, None)
    if t > 0:
        out[("gate", 0)] = (sc + "/" + _nm("Conv2d_transpose", 2 * (L - 1) + 1) + "/weights", None, "deconv", 2, None)
    return out


def _rel(a, ref, floor=0.0):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.linalg.norm(a - ref) / max(np.linalg.norm(ref), floor, 1e-30))


def _contract(kind, x, w, stride):
    if kind == "conv":
        return O.conv2d_same(x, w, stride)
    if kind == "deconv":
        return O.conv2d_transpose_same(x, w, stride)
    return O._fc(x.reshape(x.shape[0], -1), w)


def _act(v, act):
    return O.lrelu(v) if act == "lrelu" else torch.relu(v) if act == "relu" else v


def replay(model, hp, P, operand, steps=None):
    """Run after model.forward(...) + model.backward().  Returns {check: (worst error, where)} over all tensor-core blocks
    of the requested chain steps.  P: {name: torch fp64 tensor} - the weights the model holds."""
    L, T = hp["vlae_levels"], hp["mc_steps"]
    G = model.gradients()
    worst = {}
    bf16 = operand == "bf16"

    def note(check, err, where):
        if check not in worst or err > worst[check][0]:
            worst[check] = (err, where)

    def tens(t, net, idx, which):
        return torch.tensor(model.block_tensor(t, net, idx, which), dtype=torch.float64)

    gscale = max(float(np.abs(g).max()) for g in G.values())
    with oracle_mode(operand):
        for t in (range(T) if steps is None else steps):
            names = block_names(hp, t)
            dx = {}          # (net, idx) -> oracle input gradient of that block, from the device's own dy
            for (net, idx), (wname, bname, kind, stride, act) in names.items():
                where = "t%d %s%d" % (t, net, idx)
                w = P[wname].double()
                x = tens(t, net, idx, "input")
                y_gpu = tens(t, net, idx, "y")
                # ---- forward contraction
                y_ref = _contract(kind, x, w, stride).reshape(y_gpu.shape)
                note("fwd_contraction", _rel(y_gpu, y_ref), where)
                dy_gpu = tens(t, net, idx, "dy")
                if bname is not None:
                    beta = P[bname].double()
                    try:
                        res = tens(t, net, idx, "res")
                    except Exception:
                        res = None
                    flat = kind == "fc"
                    yv = (y_gpu.reshape(y_gpu.shape[0], -1) if flat else y_gpu).clone().requires_grad_(True)
                    pre = O.batch_norm(yv, beta)
                    if res is not None:
                        pre = pre + res.reshape(pre.shape)
                    out_ref = _act(pre, act)
                    # ---- batch norm + shortcut + activation (the consumer may only keep a bf16 copy: one bf16 ulp = 2^-8)
                    out_gpu = tens(t, net, idx, "out").reshape(out_ref.shape)
                    note("bn_act_out", _rel(out_gpu, out_ref.detach()), where)
                    # ---- batch-norm backward on the device's own da
                    da = tens(t, net, idx, "da").reshape(out_ref.shape)
                    # Activation branch of BORDERLINE elements.  The device evaluates the pre-activation in fp32, once in the
                    # forward kernel and once more in the backward kernel; an element within fp32 rounding of zero may land on
                    # either side in either of them (and in the oracle).  One such element moves dy and dbeta by ~1/sqrt(numel)
                    # - more than the bounds below - without any kernel being wrong.  So: elements with |pre| below a few fp32
                    # ulps of the terms that form it are borderline; for each of them the branch the device took is read off
                    # its own dy (a flipped branch shows up as a spike of rstd * da * (1 - slope) at exactly that element);
                    # every other element, the statistics, both sums, the scaling and the shortcut are compared at the bound.
                    if act is None:
                        g_pre = da
                    else:
                        slope = 0.1 if act == "lrelu" else 0.0
                        pre_d = pre.detach()
                        pos = pre_d > 0

                        def g_of(mask):
                            return da * torch.where(mask, torch.ones_like(da), torch.full_like(da, slope))

                        g_pre = g_of(pos)
                        border = pre_d.abs() < 4e-6 * (1.0 + float(pre_d.abs().mean()))
                        nb = int(border.sum())
                        if nb:
                            assert nb <= 64 + border.numel() // 20000, ("borderline pre-activations", where, nb)
                            dy0, = torch.autograd.grad(pre, yv, g_pre, retain_graph=True)
                            y2 = yv.detach().reshape(-1, yv.shape[-1])
                            rstd = (1.0 / torch.sqrt(y2.var(0, unbiased=False) + 1e-3)).expand_as(pre_d)
                            resp = ((g_of(~pos) - g_pre) * rstd)[border]
                            resid = (dy_gpu.reshape(dy0.shape) - dy0)[border]
                            flip = (resid - resp).abs() < resid.abs()
                            pos = pos.clone()
                            pos[border] = pos[border] ^ flip
                            g_pre = g_of(pos)
                    dy_ref, = torch.autograd.grad(pre, yv, g_pre)
                    note("bn_bwd_dy", _rel(dy_gpu.reshape(dy_ref.shape), dy_ref), where)
                    dbeta_ref = g_pre.reshape(-1, g_pre.shape[-1]).sum(0) if not flat else g_pre.sum(0)
                    note("dbeta", _rel(G[bname], dbeta_ref.numpy(), 1e-6 * gscale * np.sqrt(dbeta_ref.numel())), where)
                    if res is not None:
                        dx[("res", net, idx)] = g_pre
                # ---- weight gradient and input gradient from the device's own dy
                xg = x.clone().requires_grad_(True)
                wg = w.clone().requires_grad_(True)
                yy = _contract(kind, xg, wg, stride)
                gx, gw = torch.autograd.grad(yy, [xg, wg], dy_gpu.reshape(yy.shape))
                note("wgrad", _rel(G[wname], gw.numpy(), 1e-6 * gscale * np.sqrt(gw.numel())), where)
                dx[(net, idx)] = gx
            # ---- input gradients against dL/d(out) of the producing block(s)
            def cmp(net, idx, ref, where):
                da = tens(t, net, idx, "da")
                note("dgrad", _rel(da.reshape(-1), ref.reshape(-1), 1e-7 * float(da.abs().max() + 1e-30) * np.sqrt(da.numel())), where)

            for lvl in range(L - 1):
                F = hp["filter_sizes"][lvl + 1]
                if lvl >= 1:
                    cmp("tb", lvl, dx[("ta", lvl - 1)], "t%d d_c[%d] <- ta%d" % (t, lvl, lvl - 1))
                else:
                    ref = dx[("out", 0)] + (dx[("gate", 0)] if t > 0 else 0.0)
                    cmp("tb", 0, ref, "t%d d_c[0] <- out + gate" % t)
                cmp("ta", lvl, dx[("tb", lvl)][..., :F], "t%d d_dcat[%d][:F] <- tb%d" % (t, lvl, lvl))
            cmp("decfc", 0, dx[("ta", L - 2)], "t%d d_fc <- ta%d" % (t, L - 2))
            if t > 0:
                FL = hp["filter_sizes"][L]
                cmp("encfc", 0, dx[("decfc", 0)][..., :FL], "t%d d_cc[:F_L] <- decfc" % t)
                last = 2 * (L - 1)
                cmp("enc", last, dx[("encfc", 0)], "t%d d_e[%d] <- encfc" % (t, last))
                for k in range(last, 0, -1):
                    ref = dx[("enc", k)]
                    if (k - 1) % 2 == 1:                       # e[l+1] = enc[2l+1] also feeds the decoder shortcut (:1713)
                        ref = ref + dx[("res", "ta", (k - 2) // 2)]
                    cmp("enc", k - 1, ref, "t%d d_e[%d] <- enc%d" % (t, k - 1, k))
            for k in range(1, 2 * (L - 1), 2):                 # even producers only: odd ones also receive the heads' gradient
                cmp("inf", k - 1, dx[("inf", k)], "t%d d_inf[%d] <- inf%d" % (t, k - 1, k))
    return worst
