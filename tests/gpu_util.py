"""Helpers shared by the GPU parity tests (torch CUDA tensors are only the device-memory container)."""
import ctypes as C

import numpy as np
import torch

import seqvae_b200 as S
from seqvae_b200 import _cabi
from oracle import seqvae_oracle as O

TINY = dict(filter_sizes=[3, 8, 16, 16, 24, 24], vlae_latent_dims=[2, 3, 2, 2], mc_steps=2)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def dev(a, dtype=torch.float32):
    """Upload on torch's stream and wait: libsvae runs on its own stream, so the copy must be complete first."""
    t = torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda().contiguous()
    torch.cuda.synchronize()
    return t


_handle_model = {}


def op_handle():
    """A tiny model whose handle is used for the layer-level svae_op_* entry points."""
    if "m" not in _handle_model:
        ds = S.SyntheticDataset("celebA", 2, data_dims=[16, 16, 3], data_range=[-1.0, 1.0])
        _handle_model["m"] = S.SequentialVAE(ds, 2, "c_inhomog", operand_dtype="fp32", restore=False, **TINY)
    m = _handle_model["m"]
    return m, m._L, m._h


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def make_pair(netname, dims, rng, B, operand="fp32", seed=0, perturb=True, train=True, max_batch=None, **over):
    """(model on the GPU, oracle hp, oracle params) with identical weights."""
    ds = S.SyntheticDataset("x", B, data_dims=dims, data_range=list(rng))
    model = S.SequentialVAE(ds, B, netname, operand_dtype=operand, restore=False, train=train, max_batch=max_batch,
                            **over)
    hp = O.hyperparams(netname, dims, rng, **over)
    P = O.init_params(hp, seed)
    if perturb:
        g = torch.Generator().manual_seed(99)
        for k in P:
            if k.endswith("/beta") or k.endswith("/biases"):
                P[k] = (0.1 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)).float().double()
    model.set_params({k: v.numpy() for k, v in P.items()})
    return model, hp, P


def make_inputs(hp, B, seed=1):
    g = torch.Generator().manual_seed(seed)
    lo, hi = hp["range"]
    x = (torch.rand([B] + hp["data_dims"], generator=g, dtype=torch.float64) * (hi - lo) + lo).float().double()
    eps = torch.randn(hp["mc_steps"], B, hp["latent_dim"], generator=g, dtype=torch.float64).float().double()
    return x, eps


class bf16_emulation:
    """Context manager: make the oracle round the operands of exactly those contractions that libsvae's
    SVAE_OPERAND_BF16 family runs on tensor cores (queried from the library, so the two cannot drift apart)."""

    def __enter__(self):
        L = _cabi.lib()

        def pred(direction):
            return lambda kind, h, w, cin, cout, stride: bool(
                L.svae_op_tc_supported({"conv": 0, "deconv": 1, "fc": 2}[kind], h, w, cin, cout, stride, direction))

        O.OPERAND_EMULATION = dict(fwd=pred(0), dgrad=pred(1), wgrad=pred(2))
        return self

    def __exit__(self, *a):
        O.OPERAND_EMULATION = None
        return False


class no_emulation:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def oracle_mode(operand):
    return bf16_emulation() if operand == "bf16" else no_emulation()
