"""Naive numpy restatement of the TensorFlow 1.x op semantics the oracle relies on.  TEST INFRASTRUCTURE ONLY.

Pure-Python loops, small cases only.  These follow the *published* TF definitions, written
independently of ``seqvae_oracle.py`` (no torch, no shared helpers) so that the two can be checked against each other:

* ``conv2d_same``            - tf.nn.conv2d(padding='SAME'), NHWC / HWIO          (call site abstract_network.py:18)
* ``conv2d_backprop_input``  - what tf.contrib.layers.convolution2d_transpose runs (abstract_network.py:37,46,56;
                               sequential_vae.py:1720,1727): the gradient of the SAME conv w.r.t. its input
* ``batch_norm_train``       - tf.contrib.layers.batch_norm defaults in training mode (abstract_network.py:22)
* ``adam_tf``                - tf.train.AdamOptimizer update (sequential_vae.py:1267,1276)
"""
import math

import numpy as np


def same_pad_before(size, k, s):
    out = (size + s - 1) // s
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2


def conv2d_same(x, w, stride):
    """x [N,H,W,Ci], w [kh,kw,Ci,Co] -> [N,ceil(H/s),ceil(W/s),Co]; zero padding, extra pad on bottom/right."""
    N, H, W, Ci = x.shape
    kh, kw, _, Co = w.shape
    Ho, pt = same_pad_before(H, kh, stride)
    Wo, pl = same_pad_before(W, kw, stride)
    y = np.zeros((N, Ho, Wo, Co), dtype=np.float64)
    for oh in range(Ho):
        for ow in range(Wo):
            for a in range(kh):
                ih = oh * stride - pt + a
                if ih < 0 or ih >= H:
                    continue
                for b in range(kw):
                    iw = ow * stride - pl + b
                    if iw < 0 or iw >= W:
                        continue
                    y[:, oh, ow, :] += x[:, ih, iw, :] @ w[a, b]
    return y


def conv2d_backprop_input(dy, w, stride, in_hw):
    """Gradient of conv2d_same(x, w, stride) w.r.t. x.  dy [N,Ho,Wo,Co], w [kh,kw,Ci,Co] -> [N,H,W,Ci].
    convolution2d_transpose(inputs=dy, num_outputs=Ci, stride) with weights [kh,kw,Ci(out),Co(in)] is this op with
    in_hw = (Ho*stride, Wo*stride)."""
    N, Ho, Wo, Co = dy.shape
    kh, kw, Ci, _ = w.shape
    H, W = in_hw
    _, pt = same_pad_before(H, kh, stride)
    _, pl = same_pad_before(W, kw, stride)
    dx = np.zeros((N, H, W, Ci), dtype=np.float64)
    for oh in range(Ho):
        for ow in range(Wo):
            for a in range(kh):
                ih = oh * stride - pt + a
                if ih < 0 or ih >= H:
                    continue
                for b in range(kw):
                    iw = ow * stride - pl + b
                    if iw < 0 or iw >= W:
                        continue
                    dx[:, ih, iw, :] += dy[:, oh, ow, :] @ w[a, b].T
    return dx


def batch_norm_train(x, beta, eps=1e-3):
    axes = tuple(range(x.ndim - 1))
    mean = x.mean(axis=axes, keepdims=True)
    var = x.var(axis=axes, keepdims=True)          # biased
    return (x - mean) / np.sqrt(var + eps) + beta


def adam_tf(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v
