"""SequentialVAE - host-side mirror of the reference class surface (reference sequential_vae.py:28-1455) over libsvae.

Kept from the reference so that ``main.py`` / ``trainer.py`` style callers drop in unchanged:
``SequentialVAE(dataset, batch_size, name, logger, version, base_dir, num_gpus)`` (:81), ``train`` (:1341), ``test``
(:1381), ``generate_mc_samples`` (:1397), ``training_mc_samples`` (:1434), ``save_network`` / ``init_network``
(abstract_network.py:124-152) and the public attributes ``name, iteration, learning_rate, mc_steps, latent_dim,
batch_size, data_dims``.  Everything the reference did inside ``Session.run`` happens inside libsvae.so; this file only
schedules (lr decay, KL warm-up :1351-1357), marshals buffers and returns numpy arrays.

Additions that the TF session made implicit: ``forward`` / ``backward`` / ``gradients`` probes (TF lets callers fetch any
tensor; parity tests need mu, sigma, x_t, per-step ELBO terms and gradients) and explicit ``eps`` / ``z`` injection
(the reference draws them in-graph, unseeded, :1023,1417; parity needs identical noise).
"""
import ctypes as C
import logging
import math
import os
from collections import OrderedDict

import numpy as np

from . import _cabi
from .checkpoint import adam_t_from_beta_powers, beta_powers, tf_checkpoint_layout
from .config import hyperparams, to_cabi_config


def _is_cuda_tensor(x):
    return hasattr(x, "is_cuda") and bool(getattr(x, "is_cuda"))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class SequentialVAE:
    def __init__(self, dataset, batch_size, name, logger=None, version=0, base_dir=None, num_gpus=1, *, device=0,
                 operand_dtype=None, train=True, max_batch=None, seed=0, restore=True, **overrides):
        # --- attribute block (sequential_vae.py:195-258, abstract_network.py:85-107)
        self.dataset = dataset
        self.batch_size = batch_size
        self.data_dims = list(dataset.data_dims)
        self.LOG = logger if logger is not None else logging.getLogger("seqvae_b200")
        self.num_gpus = num_gpus
        self.name = name
        self.version = version
        self.base_dir = base_dir if base_dir is not None else "models/" + name + "_v" + str(version)
        self.iteration = 0
        try:
            hp = hyperparams(name, self.data_dims, dataset.range, **overrides)
        except KeyError:
            self.LOG.error("Unknown network name %s" % name)      # sequential_vae.py:860-862 (exit(-1) there)
            raise
        self.hp = hp
        self.vlae_levels = hp["vlae_levels"]
        self.vlae_latent_dims = list(hp["vlae_latent_dims"])
        self.image_sizes = list(hp["image_sizes"])
        self.filter_sizes = list(hp["filter_sizes"])
        self.mc_steps = hp["mc_steps"]
        self.latent_dim = hp["latent_dim"]
        self.learning_rate = hp["learning_rate"]
        self.learning_rate_decay = hp["learning_rate_decay"]
        self.reg_coeff_rate = hp["reg_coeff_rate"]
        self.share_theta_weights = bool(hp["share_theta_weights"])        # sequential_vae.py:213-214
        self.share_phi_weights = bool(hp["share_phi_weights"])
        self.add_noise_to_chain = bool(hp["add_noise_to_chain"])          # :233
        self.noise_stddevs = list(hp["noise_stddevs"])                    # :239
        self.save_freq = hp["save_freq"]
        # A main.py-style caller (SequentialVAE(dataset, batch_size=..., name=...)) gets the production kernel family: bf16
        # operands on the tcgen05 tensor cores with fp32 accumulation.  operand_dtype="fp32" (or SVAE_OPERAND=fp32) selects the
        # strict-parity fp32 SIMT family.
        if operand_dtype is None:
            operand_dtype = os.environ.get("SVAE_OPERAND", "bf16")
        if operand_dtype not in ("fp32", "bf16"):
            raise ValueError("operand_dtype must be 'fp32' or 'bf16', got %r" % (operand_dtype,))
        self.operand_dtype = operand_dtype
        self.device = device
        self.max_batch = int(max_batch if max_batch is not None else batch_size)
        self.last_losses = None
        self._h = None
        self._stream = None

        # --- build the static plan + device state (construct_network / init_network equivalents)
        cfg = to_cabi_config(hp, self.max_batch, train, operand_dtype)
        self._cfg = cfg
        L = _cabi.lib()
        h = C.c_void_p()
        rc = L.svae_create(C.byref(cfg), int(device), C.byref(h))
        if rc != 0:
            _cabi.check(None, rc)
        self._h = h
        self._L = L
        self._params = self._read_param_table()
        self.init_network(seed=seed, restore=restore)
        self.log_tf_variables()

    # ------------------------------------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._L.svae_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        return _cabi.check(self._h, rc)

    # ------------------------------------------------------------------------------------------------ parameters
    def _read_param_table(self):
        out = []
        n = self._L.svae_param_count(self._h)
        for i in range(n):
            info = _cabi.ParamInfo()
            self._chk(self._L.svae_param_info_get(self._h, i, C.byref(info)))
            out.append(dict(index=i, name=info.name.decode(), shape=tuple(info.shape[:info.ndim]), numel=info.numel,
                            offset=info.offset, step=info.step, flags=info.flags))
        return out

    @property
    def param_table(self):
        """[{name, shape, flags, step, ...}] in tf.trainable_variables() order."""
        return self._params

    def param_slices(self, name):
        """Arena element offsets of the per-chain-step slices of variable ``name`` (one for an unshared variable; one per
        step that uses it for a shared variable of a homogeneous chain, sequential_vae.py:213-214)."""
        idx = {p["name"]: p["index"] for p in self._params}[name]
        offs = (C.c_int64 * 64)()
        n = self._L.svae_param_slices(self._h, idx, offs, 64)
        if n < 0:
            self._chk(n)
        return [int(offs[i]) for i in range(n)]

    def read_arena(self, which="param"):
        """The whole flat fp32 arena ("param", "grad", "adam_m", "adam_v") as a numpy array (offsets: ``param_table`` /
        ``param_slices``)."""
        n = int(self._L.svae_arena_numel(self._h))
        buf = np.empty(n, np.float32)
        self._chk(self._L.svae_arena_read(self._h, {"param": 0, "grad": 1, "adam_m": 2, "adam_v": 3}[which], 0, n,
                                          buf.ctypes.data_as(C.c_void_p)))
        return buf

    def init_network(self, seed=0, restore=True):
        """abstract_network.py:139-152: restore the checkpoint if one exists, else reference initialisers
        (N(0,0.02) for *_bn_* blocks, xavier-uniform for heads / output deconvs, zeros for biases and betas)."""
        ckpt = os.path.join(self.base_dir, self.name + ".npz")
        if restore and os.path.exists(ckpt):
            try:
                self.load_network(ckpt)
                self.LOG.info("Restored network from %s" % ckpt)
                return
            except (OSError, KeyError, ValueError, EOFError) as e:
                # abstract_network.py:146-150: an unreadable / incompatible checkpoint -> warn and re-initialise.  Anything
                # else (a library error half-way through the upload, ...) propagates: silently training from scratch on top
                # of a partially restored handle would discard the checkpoint without telling anyone.
                self.LOG.warning("Could not restore %s (%s); re-initialising" % (ckpt, e))
        rng = np.random.default_rng(seed)
        for p in self._params:
            name, shape = p["name"], p["shape"]
            if name.endswith("/weights"):
                if p["flags"] & _cabi.PF_XAVIER:
                    if len(shape) == 2:
                        fan = shape[0] + shape[1]
                    else:
                        fan = shape[0] * shape[1] * (shape[2] + shape[3])
                    lim = math.sqrt(6.0 / fan)
                    v = rng.uniform(-lim, lim, size=shape)
                else:
                    v = rng.normal(0.0, 0.02, size=shape)
            else:
                v = np.zeros(shape)
            self._set_param(p["index"], v)

    def _set_param(self, idx, value):
        v = _f32(value)
        assert v.size == self._params[idx]["numel"], (self._params[idx]["name"], v.shape)
        self._chk(self._L.svae_param_set(self._h, idx, v.ctypes.data_as(C.c_void_p)))

    def set_params(self, values):
        """values: {tf variable name: array}; names not present keep their current value."""
        by_name = {p["name"]: p for p in self._params}
        for k, v in values.items():
            self._set_param(by_name[k]["index"], np.asarray(v).reshape(by_name[k]["shape"]))

    def _fetch(self, fn, live_only=False):
        out = OrderedDict()
        for p in self._params:
            if live_only and p["flags"] & (_cabi.PF_INERT | _cabi.PF_DEAD):
                continue
            buf = np.empty(p["shape"], dtype=np.float32)
            self._chk(fn(self._h, p["index"], buf.ctypes.data_as(C.c_void_p)))
            out[p["name"]] = buf
        return out

    def get_params(self, live_only=False):
        return self._fetch(self._L.svae_param_get, live_only)

    def gradients(self, live_only=False):
        """Gradients of the last ``backward`` (optimizer.compute_gradients, sequential_vae.py:1273)."""
        return self._fetch(self._L.svae_grad_get, live_only)

    def log_tf_variables(self):
        """sequential_vae.py:1326-1335."""
        self.LOG.debug("Printing out all variable names constructed (for debugging).")
        for i, p in enumerate(self._params):
            self.LOG.debug("(%dth variable) %s" % (i, p["name"]))

    # ------------------------------------------------------------------------------------------------ checkpoint
    def save_network(self):
        """abstract_network.py:124-135 (tf.train.Saver) -> one .npz holding exactly the variable set the TF Saver writes
        (``checkpoint.tf_checkpoint_layout``: trainable variables, batch-norm moving statistics, Adam slots, beta powers)
        plus the host-side schedule state the reference forgets (``__iteration``, ``__learning_rate``, ``__adam_t``).  An
        existing file is moved to ``<models>/old`` first, like the reference."""
        os.makedirs(self.base_dir, exist_ok=True)
        path = os.path.join(self.base_dir, self.name + ".npz")
        train = bool(self._cfg.train_capacity)
        values = self.get_params()
        slots = {}
        if train:
            for p in self._params:
                m = np.empty(p["shape"], np.float32)
                v = np.empty(p["shape"], np.float32)
                self._chk(self._L.svae_adam_get(self._h, p["index"], m.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p)))
                slots[p["name"]] = (m, v)
        adam_t = int(self._L.svae_adam_step_count(self._h)) if train else 0
        b1p, b2p = beta_powers(adam_t, self._cfg.adam_beta1, self._cfg.adam_beta2)
        blob = {}
        for name, shape, kind, src in tf_checkpoint_layout(self._params, train):
            if kind == "param":
                blob[name] = values[src]
            elif kind == "bn_moving_mean":
                blob[name] = np.zeros(shape, np.float32)          # never updated by the reference
            elif kind == "bn_moving_variance":
                blob[name] = np.ones(shape, np.float32)
            elif kind == "adam_m":
                blob[name] = slots[src][0]
            elif kind == "adam_v":
                blob[name] = slots[src][1]
            elif kind == "beta1_power":
                blob[name] = np.float32(b1p)
            elif kind == "beta2_power":
                blob[name] = np.float32(b2p)
        if train:
            blob["__adam_t"] = np.int64(adam_t)
        blob["__iteration"] = np.int64(self.iteration)
        blob["__learning_rate"] = np.float64(self.learning_rate)
        # write next to the target, then swap: a crash mid-save leaves the previous checkpoint in place
        tmp = path + ".tmp.npz"
        np.savez(tmp, **blob)
        if os.path.exists(path):
            old = os.path.join(os.path.dirname(self.base_dir) or ".", "old")
            os.makedirs(old, exist_ok=True)
            os.replace(path, os.path.join(old, self.name + "_v" + str(self.version) + ".npz"))
        os.replace(tmp, path)
        self.LOG.info("Saved network to %s" % path)
        return path

    def load_network(self, path):
        """Restore from ``save_network``'s file (or any .npz keyed by the TF variable names, e.g. a converted TF checkpoint):
        every trainable variable must be present (tf.train.Saver.restore fails on a missing key too); Adam slots, beta
        powers and the schedule keys are optional."""
        blob = np.load(path)
        missing = [p["name"] for p in self._params if p["name"] not in blob.files]
        if missing:
            raise KeyError("checkpoint %s lacks %d variables, e.g. %s" % (path, len(missing), missing[0]))
        self.set_params({p["name"]: blob[p["name"]] for p in self._params})
        if self._cfg.train_capacity:
            have_slots = False
            for p in self._params:
                if p["name"] + "/Adam" in blob.files:
                    m, v = _f32(blob[p["name"] + "/Adam"]), _f32(blob[p["name"] + "/Adam_1"])
                    have_slots = True
                else:                                       # dead-branch variables have no slots in a TF checkpoint
                    m = v = np.zeros(p["shape"], np.float32)
                self._chk(self._L.svae_adam_set(self._h, p["index"], m.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p)))
            if "__adam_t" in blob.files:
                self._chk(self._L.svae_adam_set_step_count(self._h, int(blob["__adam_t"])))
            elif have_slots and ("beta2_power" in blob.files or "beta1_power" in blob.files):
                self._chk(self._L.svae_adam_set_step_count(self._h, adam_t_from_beta_powers(
                    blob["beta1_power"] if "beta1_power" in blob.files else None,
                    blob["beta2_power"] if "beta2_power" in blob.files else None,
                    self._cfg.adam_beta1, self._cfg.adam_beta2)))
        if "__iteration" in blob.files:
            self.iteration = int(blob["__iteration"])
            self.learning_rate = float(blob["__learning_rate"])

    # ------------------------------------------------------------------------------------------------ buffers
    def use_torch_stream(self, stream=None):
        """Order all device work on a torch CUDA stream (so that torch.cuda.Event timing and tensor lifetimes are
        consistent).  ``None`` -> torch's current stream."""
        import torch

        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._stream = s
        self._chk(self._L.svae_set_stream(self._h, C.c_void_p(s.cuda_stream)))

    def sync(self):
        self._chk(self._L.svae_sync(self._h))

    def _losses(self, ls):
        T = self.mc_steps
        return dict(loss=float(ls.total), final_loss=float(ls.final_recon), recon=[float(ls.recon[t]) for t in range(T)],
                    kl=[float(ls.kl[t]) for t in range(T)])

    def _check_batch(self, x):
        shp = tuple(x.shape)
        if len(shp) != 4 or list(shp[1:]) != self.data_dims:
            raise ValueError("expected a [B,%d,%d,%d] batch, got %s" % (*self.data_dims, shp))   # TF raises on shape mismatch
        if shp[0] < 1 or shp[0] > self.max_batch:
            raise ValueError("batch %d exceeds max_batch %d" % (shp[0], self.max_batch))
        return shp[0]

    def _device_args(self, x, tgt, eps):
        """Device-pointer path: libsvae reads raw ``data_ptr()`` values, so the tensors must be what the C ABI documents
        (dense float32 on this handle's device, target shaped like the input, eps [T,B,Z]) and the work that PRODUCED them
        must be ordered before the library's stream: when the handle runs on its own stream (no ``use_torch_stream``, or a
        different torch stream is current) the producing stream is drained first.  Converted copies are kept alive until
        the next call - the library's stream may still be reading them when this method returns."""
        import torch

        B = self._check_batch(x)
        if not _is_cuda_tensor(tgt) or tuple(tgt.shape) != tuple(x.shape):
            raise ValueError("batch_target must be a CUDA tensor shaped like input_batch, got %s" % (tuple(getattr(tgt, "shape", ())),))
        if eps is not None and (not _is_cuda_tensor(eps) or tuple(eps.shape) != (self.mc_steps, B, self.latent_dim)):
            raise ValueError("eps must be a CUDA tensor of shape [T,B,Z] = %s" % ((self.mc_steps, B, self.latent_dim),))
        dev = torch.device("cuda", self.device)
        for t in (x, tgt) + (() if eps is None else (eps,)):
            if t.device != dev:
                raise ValueError("tensor lives on %s, the handle on %s" % (t.device, dev))
        conv = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous()) else t.contiguous().float()
        same = tgt is x
        x = conv(x)
        tgt = x if same else conv(tgt)
        eps = None if eps is None else conv(eps)
        cur = torch.cuda.current_stream(dev)
        if self._stream is None or self._stream.cuda_stream != cur.cuda_stream:
            cur.synchronize()
        self._keep, self._keep_prev = (x, tgt, eps), getattr(self, "_keep", None)
        return x, tgt, eps, B

    # ------------------------------------------------------------------------------------------------ run wrappers
    def train(self, input_batch, batch_target, eps=None, seed=None):
        """ONE training update (sequential_vae.py:1341-1375): schedules, forward + backward + clipped Adam, periodic
        save; returns final_loss / H / W.  numpy inputs take the host-buffer path (H2D inside the call, like
        feed_dict); torch CUDA tensors take the device-pointer path."""
        self.iteration += 1
        self.learning_rate *= self.learning_rate_decay
        reg = 1 - math.exp(-self.iteration / self.reg_coeff_rate)                    # :1357
        seed = int(self.iteration if seed is None else seed)
        ls = _cabi.Losses()
        if _is_cuda_tensor(input_batch):
            x, tgt, e, B = self._device_args(input_batch, batch_target, eps)
            self._chk(self._L.svae_train_step(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(tgt.data_ptr()), B,
                                              C.c_void_p(e.data_ptr()) if e is not None else None, seed,
                                              float(self.learning_rate), float(reg)))
            self._chk(self._L.svae_read_losses(self._h, C.byref(ls)))
        else:
            x, tgt = _f32(input_batch), _f32(batch_target)
            B = self._check_batch(x)
            e = None if eps is None else _f32(eps)
            self._chk(self._L.svae_train_step_host(
                self._h, x.ctypes.data_as(C.c_void_p), tgt.ctypes.data_as(C.c_void_p), B,
                e.ctypes.data_as(C.c_void_p) if e is not None else None, seed, float(self.learning_rate), float(reg),
                C.byref(ls)))
        self.last_losses = self._losses(ls)
        if self.iteration % self.save_freq == 0:                                     # :1368-1369
            self.save_network()
        return self.last_losses["final_loss"] / self.data_dims[0] / self.data_dims[1]   # :1375

    # --- denoising corruption on the device (NoisyTrainer.apply_noise, trainer.py:56-78; constants trainer.py:16-18)
    def apply_noise(self, original, pepper_prob=0.1, salt_prob=0.1, gaussian_noise_scale=0.1, seed=0, return_draws=False):
        """clip(original * Bernoulli(1 - pepper) + Bernoulli(salt) + N(0, scale), *dataset.range) with device-side Philox
        draws keyed by ``seed``.  numpy in -> numpy out; a torch CUDA tensor is corrupted into a new CUDA tensor without
        leaving the device.  return_draws: also return the [3, ...] keep / salt / Gaussian fields that were used."""
        lo, hi = float(self.hp["range"][0]), float(self.hp["range"][1])
        args = (float(pepper_prob), float(salt_prob), float(gaussian_noise_scale), lo, hi, int(seed))
        if _is_cuda_tensor(original):
            import torch

            x = original.contiguous().float()
            out = torch.empty_like(x)
            draws = torch.empty((3,) + tuple(x.shape), dtype=torch.float32, device=x.device) if return_draws else None
            torch.cuda.current_stream(x.device).synchronize()      # libsvae runs on its own stream
            self._chk(self._L.svae_apply_noise(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), x.numel(),
                                               *args, C.c_void_p(draws.data_ptr()) if draws is not None else None))
            self.sync()
            return (out, draws) if return_draws else out
        x = _f32(original)
        out = np.empty_like(x)
        draws = np.empty((3,) + x.shape, np.float32) if return_draws else None
        self._chk(self._L.svae_apply_noise_host(self._h, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                                x.size, *args,
                                                draws.ctypes.data_as(C.c_void_p) if draws is not None else None))
        return (out, draws) if return_draws else out

    def train_denoise(self, batch_target, pepper_prob=0.1, salt_prob=0.1, gaussian_noise_scale=0.1, eps=None, seed=None,
                      noise_seed=None, return_input=False):
        """trainer.py:100-104 with --denoise_train as ONE call: ``train(apply_noise(batch), batch)`` where the corruption
        happens on the device - the clean batch is uploaded once (half the reference's host->device bytes, no host RNG)."""
        self.iteration += 1
        self.learning_rate *= self.learning_rate_decay
        reg = 1 - math.exp(-self.iteration / self.reg_coeff_rate)                    # :1357
        seed = int(self.iteration if seed is None else seed)
        noise_seed = int(self.iteration if noise_seed is None else noise_seed)
        x = _f32(batch_target)
        B = self._check_batch(x)
        e = None if eps is None else _f32(eps)
        noisy = np.empty_like(x) if return_input else None
        ls = _cabi.Losses()
        self._chk(self._L.svae_train_step_host_denoise(
            self._h, x.ctypes.data_as(C.c_void_p), B, e.ctypes.data_as(C.c_void_p) if e is not None else None, seed,
            float(self.learning_rate), float(reg), float(pepper_prob), float(salt_prob), float(gaussian_noise_scale),
            noise_seed, noisy.ctypes.data_as(C.c_void_p) if noisy is not None else None, C.byref(ls)))
        self.last_losses = self._losses(ls)
        if self.iteration % self.save_freq == 0:
            self.save_network()
        r = self.last_losses["final_loss"] / self.data_dims[0] / self.data_dims[1]
        return (r, noisy) if return_input else r

    def train_async(self, x_dev, tgt_dev, eps_dev=None, seed=None):
        """Device-resident, non-blocking variant of ``train`` for benchmarking: no loss read-back, no sync."""
        self.iteration += 1
        self.learning_rate *= self.learning_rate_decay
        reg = 1 - math.exp(-self.iteration / self.reg_coeff_rate)
        x_dev, tgt_dev, eps_dev, B = self._device_args(x_dev, tgt_dev, eps_dev)
        self._chk(self._L.svae_train_step(self._h, C.c_void_p(x_dev.data_ptr()), C.c_void_p(tgt_dev.data_ptr()), B,
                                          C.c_void_p(eps_dev.data_ptr()) if eps_dev is not None else None,
                                          int(self.iteration if seed is None else seed), float(self.learning_rate),
                                          float(reg)))

    def forward(self, input_batch, batch_target=None, eps=None, reg_coeff=1.0, seed=0):
        """Training-mode chain probe: per-step mu, sigma, x_t and ELBO terms (what sess.run on training_mles /
        latents / the loss summaries would return)."""
        x = _f32(input_batch)
        tgt = x if batch_target is None else _f32(batch_target)
        B = self._check_batch(x)
        T, Z = self.mc_steps, self.latent_dim
        e = None if eps is None else _f32(eps)
        mu = np.empty((T, B, Z), np.float32)
        sd = np.empty((T, B, Z), np.float32)
        xs = np.empty([T, B] + self.data_dims, np.float32)
        ls = _cabi.Losses()
        self._chk(self._L.svae_forward_host(
            self._h, x.ctypes.data_as(C.c_void_p), tgt.ctypes.data_as(C.c_void_p), B,
            e.ctypes.data_as(C.c_void_p) if e is not None else None, int(seed), float(reg_coeff),
            mu.ctypes.data_as(C.c_void_p), sd.ctypes.data_as(C.c_void_p), xs.ctypes.data_as(C.c_void_p), None,
            C.byref(ls)))
        out = self._losses(ls)
        out.update(mu=mu, sigma=sd, x=xs)
        self.last_losses = out
        return out

    def backward(self):
        """Reverse-mode through the whole chain for the last ``forward``; read results with ``gradients()``."""
        self._chk(self._L.svae_backward(self._h))
        self.sync()

    _NETS = dict(inf=0, enc=1, encfc=2, lat=3, decfc=4, ta=5, tb=6, out=7, gate=8)
    _WHICH = dict(input=0, y=1, out=2, da=3, dy=4, res=5)

    def block_tensor(self, step, net, index, which):
        """Parity probe (TF lets callers ``sess.run`` any tensor of the graph): one tensor of one block of the last
        ``forward`` / ``backward`` as the kernels consumed or produced it, as a dense [B,H,W,C] float32 array.
        net: inf | enc | encfc | lat | decfc | ta | tb | out | gate; which: input | y | out | da | dy | res
        (include/svae.h, svae_debug_block_tensor)."""
        dims = (C.c_int32 * 4)()
        args = (self._h, int(step), self._NETS[net], int(index), self._WHICH[which])
        self._chk(self._L.svae_debug_block_tensor(*args, None, 0, dims))
        out = np.empty([int(d) for d in dims], np.float32)
        self._chk(self._L.svae_debug_block_tensor(*args, out.ctypes.data_as(C.c_void_p), out.size, dims))
        return out

    def adam_step(self, learning_rate=None):
        self._chk(self._L.svae_adam_step(self._h, float(self.learning_rate if learning_rate is None else learning_rate)))

    def test(self, input_batch, eps=None, seed=0):
        """sequential_vae.py:1381-1391: training-mode chain (reg_coeff default 1.0), returns the final mle x_T."""
        x = _f32(input_batch)
        B = self._check_batch(x)
        e = None if eps is None else _f32(eps)
        last = np.empty([B] + self.data_dims, np.float32)
        self._chk(self._L.svae_forward_host(
            self._h, x.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), B,
            e.ctypes.data_as(C.c_void_p) if e is not None else None, int(seed), 1.0, None, None, None,
            last.ctypes.data_as(C.c_void_p), None))
        return last

    def training_mc_samples(self, input_batch, eps=None, seed=0):
        """sequential_vae.py:1434-1455: list of the T training-mode samples; with add_noise_to_chain the T mles followed by
        the T samples (:1450-1451)."""
        out = self.forward(input_batch, None, eps, 1.0, seed)
        mles = [out["x"][t] for t in range(self.mc_steps)]
        if not self.add_noise_to_chain:
            return mles
        smp = self.chain_samples(out["x"].shape[1])
        return mles + [smp[t] for t in range(self.mc_steps)]

    # ---- chain noise (add_noise_to_chain, sequential_vae.py:1088-1091) ------------------------------------------------
    def set_chain_noise(self, noise):
        """Inject the tf.random_normal(image_batch_shape) draws of every chain step ([T,B,H,W,C]) for the following calls with
        that batch size (parity tests feed the oracle the same values); None: counter-based Philox draws in the kernel."""
        if noise is None:
            self._chk(self._L.svae_set_chain_noise_host(self._h, None, 0))
            return
        n = _f32(noise)
        if n.ndim != 5 or list(n.shape[2:]) != list(self.data_dims) or n.shape[0] != self.mc_steps:
            raise ValueError("chain noise must be [T,B,H,W,C]")
        self._chk(self._L.svae_set_chain_noise_host(self._h, n.ctypes.data_as(C.c_void_p), int(n.shape[1])))

    def chain_samples(self, batch):
        """The samples x_t + noise of the last forward / generation, [T,B,H,W,C] (training_samples / generative_samples)."""
        out = np.empty([self.mc_steps, int(batch)] + self.data_dims, np.float32)
        self._chk(self._L.svae_read_chain_samples_host(self._h, out.ctypes.data_as(C.c_void_p), int(batch)))
        return out

    def generate_mc_samples(self, input_batch, batch_size=None, z=None, seed=0):
        """sequential_vae.py:1397-1428: generation-mode chain.  Returns T+1 arrays; the first is the uniform-noise x_0
        the reference prepends (:947-952).  Only the batch dimension of ``input_batch`` is used, as in the reference.
        z: optional [T,B,Z] latents (the reference draws np.random.normal on the host, :1417-1418); None draws them
        on the device with Philox from ``seed``."""
        if batch_size is None:
            batch_size = self.batch_size if input_batch is None else int(np.shape(input_batch)[0])
        B = int(batch_size)
        if B < 1 or B > self.max_batch:
            raise ValueError("batch %d exceeds max_batch %d" % (B, self.max_batch))
        zz = None if z is None else _f32(z)
        if zz is not None and zz.shape != (self.mc_steps, B, self.latent_dim):
            raise ValueError("z must be [T,B,Z]")
        out = np.empty([self.mc_steps, B] + self.data_dims, np.float32)
        self._chk(self._L.svae_generate_host(self._h, B, zz.ctypes.data_as(C.c_void_p) if zz is not None else None,
                                             int(seed), out.ctypes.data_as(C.c_void_p)))
        x0 = np.random.default_rng(seed).uniform(0.0, 1.0, size=[B] + self.data_dims).astype(np.float32)
        if self.add_noise_to_chain:      # generative_mles + generative_samples (:1424-1425): T mles, x_0, T samples
            smp = self.chain_samples(B)
            return [out[t] for t in range(self.mc_steps)] + [x0] + [smp[t] for t in range(self.mc_steps)]
        return [x0] + [out[t] for t in range(self.mc_steps)]

    def generate_async(self, batch, out_dev, z_dev=None, seed=0):
        """Device-resident generation for benchmarking: out_dev [T,B,H,W,C] torch CUDA tensor."""
        self._chk(self._L.svae_generate(self._h, int(batch), C.c_void_p(z_dev.data_ptr()) if z_dev is not None else None,
                                        int(seed), C.c_void_p(out_dev.data_ptr())))

    def visualize(self, epoch, *a, **k):
        """sequential_vae.py:1461-1531 writes image grids with scipy.misc / matplotlib: out of scope."""
        self.LOG.debug("visualize(%s): not part of the B200 hot path" % epoch)

    # ------------------------------------------------------------------------------------------------ introspection
    @property
    def launch_count(self):
        return int(self._L.svae_launch_count(self._h))

    def profile(self, on):
        """Bracket every kernel launch with CUDA events (per-kernel-class totals via ``profile_read``)."""
        self._chk(self._L.svae_profile_enable(self._h, int(bool(on))))

    def profile_read(self):
        """{class name: dict(launches, ms, flops, bytes)} since the last read; synchronises and resets."""
        arr = (_cabi.KernelStats * 32)()
        n = self._L.svae_profile_read(self._h, arr, 32)
        if n < 0:
            self._chk(n)
        return {arr[i].name.decode(): dict(launches=int(arr[i].launches), ms=float(arr[i].total_ms),
                                          flops=float(arr[i].flops), bytes=float(arr[i].bytes))
                for i in range(n) if arr[i].launches}

    @property
    def tc_layers(self):
        return int(self._L.svae_tc_layers(self._h))

    @property
    def activation_bytes(self):
        return int(self._L.svae_activation_bytes(self._h))
