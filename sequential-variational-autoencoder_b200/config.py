"""Hyper-parameter table of the benchmarked configurations.

Mirrors the attribute block of ``SequentialVAE.__init__`` (reference sequential_vae.py:197-258) and the netname rows
that are on the hot path: ``c_inhomog`` (:727), ``sequential_vae_celebA_inhomog`` (:671),
``sequential_vae_lsun`` (:712-714), ``m_inhomog`` (:842-848), plus the homogeneous (weight-shared) chains and the rows
that only move knobs this path already has.  The reference's remaining netnames select ablations that
are out of scope (InfoMax, chain noise, predicted stddev, early stopping, flat / PixelCNN decoders)."""
import math

NETNAMES = {
    "c_inhomog": {},
    "sequential_vae_celebA_inhomog": {},
    "sequential_vae_lsun": {"vlae_latent_dims": [20, 30, 30, 30]},
    "m_inhomog": {"vlae_levels": 3, "vlae_latent_dims": [2, 2, 2], "image_sizes": [32, 16, 8, 4],
                  "filter_sizes": [None, 64, 128, 192, 256], "mc_steps": 5},
    # homogeneous (weight-shared) chains, sequential_vae.py:213-214, and rows that only move existing knobs
    "sequential_vae_celebA_homog": {"share_theta_weights": True, "share_phi_weights": True},               # :675
    "sequential_vae_celebA_homog_fixed_length": {"share_theta_weights": True},                             # :709
    "c_homog": {"share_theta_weights": True, "share_phi_weights": True, "mc_steps": 25},                   # :730
    "c_homog_v1": {"vlae_latent_dims": [12, 12, 12, 12], "filter_sizes": [None, 16, 32, 64, 128, 384],     # :316
                   "share_theta_weights": True, "share_phi_weights": True},
    "c_homog_one_step": {"vlae_latent_dims": [12, 12, 12, 12], "filter_sizes": [None, 16, 32, 64, 128, 384],   # :281
                         "share_theta_weights": True, "share_phi_weights": True, "mc_steps": 1},
    "s_homog_one_step": {"vlae_latent_dims": [12, 12, 12, 12], "share_theta_weights": True,                # :301
                         "share_phi_weights": True, "mc_steps": 1},
    "vlae_celebA": {"mc_steps": 1, "vlae_latent_dims": [16, 16, 16, 16]},                                  # :704
    # chain noise with the fixed per-step stddevs of :239 (first slice): x_t + reg * stddev_t * N(0,I) feeds step t+1
    "c_sample_images": {"add_noise_to_chain": True},                                                        # :761
    "c_homog_sample_images": {"share_theta_weights": True, "share_phi_weights": True, "add_noise_to_chain": True},   # :764
    "sequential_vae_lsun_final": {"vlae_latent_dims": [20, 30, 30, 30], "intermediate_reconstruction": False},   # :721
}


def hyperparams(name, data_dims, data_range, **overrides):
    """Defaults (sequential_vae.py:201-258) + netname row + overrides.  Unknown names raise KeyError (the reference
    logs an error and calls exit(-1), sequential_vae.py:860-862)."""
    if name not in NETNAMES:
        raise KeyError("Unknown network name %s" % name)
    D, C = int(data_dims[0]), int(data_dims[-1])
    hp = dict(
        name=name, data_dims=[int(d) for d in data_dims], range=[float(data_range[0]), float(data_range[1])],
        vlae_levels=4, vlae_latent_dims=[3, 3, 3, 3],
        image_sizes=[D, D // 2, D // 4, D // 8, D // 16],
        filter_sizes=[C, 32, 64, 128, 384, 512],
        mc_steps=8, intermediate_reconstruction=True, first_step_loss_coeff=1.0,
        latent_mean_clip=math.inf, latent_prior_stddev=1.0, max_highway_ratio=1.0, min_highway_ratio=0.0,
        learning_rate=0.0002, learning_rate_decay=1.0, reg_coeff_rate=5000.0, save_freq=2000,
        clip_grads=True, clip_grad_value=10.0,
        share_theta_weights=False, share_phi_weights=False,
        add_noise_to_chain=False, noise_stddevs=[0.5 ** 1, 0.5 ** 2, 0.5 ** 3, 0.5 ** 4, 0.5 ** 5, 0.5 ** 6, 0.5 ** 7, 0],   # :233,239
    )
    # sequential_vae.py:224 evaluates range(self.mc_steps) while mc_steps still holds its default 8, BEFORE the netname rows
    # run: a row that lengthens the chain (c_homog: 25 steps, :733) keeps the KL term on steps 0..7 only
    hp["regularized_steps"] = list(range(hp["mc_steps"]))
    row = dict(NETNAMES[name])
    if "filter_sizes" in row:
        row["filter_sizes"] = [C if f is None else f for f in row["filter_sizes"]]
    hp.update(row)
    hp.update(overrides)
    hp["latent_dim"] = int(sum(hp["vlae_latent_dims"]))
    hp["regularized_steps"] = [int(t) for t in hp["regularized_steps"] if 0 <= int(t) < hp["mc_steps"]]
    L = hp["vlae_levels"]
    if "image_sizes" not in row and "image_sizes" not in overrides:
        hp["image_sizes"] = [D >> i for i in range(L + 1)]
    if hp["image_sizes"] != [D >> i for i in range(L + 1)]:
        # the reference aborts when image_sizes disagree with the conv stack (sequential_vae.py:1617-1627)
        raise ValueError("self.image_sizes and image_sizes in inference/generative networks don't match")
    if len(hp["filter_sizes"]) != L + 2 or len(hp["vlae_latent_dims"]) != L:
        raise ValueError("filter_sizes needs vlae_levels+2 entries and vlae_latent_dims vlae_levels entries")
    if hp["add_noise_to_chain"] and len(hp["noise_stddevs"]) != hp["mc_steps"]:
        # "an array which must be exactly of length self.mc_steps" (sequential_vae.py:151): the graph indexes it per step (:1736)
        raise ValueError("noise_stddevs needs exactly mc_steps entries")
    return hp


def to_cabi_config(hp, max_batch, train=True, operand_dtype="fp32"):
    """Fill the POD ``svae_config`` (include/svae.h) from a hyper-parameter dict."""
    from . import _cabi

    cfg = _cabi.Config()
    cfg.height, cfg.width, cfg.channels = hp["data_dims"]
    cfg.levels = hp["vlae_levels"]
    for i, v in enumerate(hp["vlae_latent_dims"]):
        cfg.latent_dims[i] = v
    for i, v in enumerate(hp["filter_sizes"]):
        cfg.filter_sizes[i] = v
    cfg.mc_steps = hp["mc_steps"]
    cfg.intermediate_reconstruction = int(hp["intermediate_reconstruction"])
    cfg.regularized_mask = sum(1 << t for t in hp["regularized_steps"])
    cfg.first_step_loss_coeff = hp["first_step_loss_coeff"]
    cfg.latent_mean_clip = hp["latent_mean_clip"]
    cfg.prior_stddev = hp["latent_prior_stddev"]
    cfg.min_highway, cfg.max_highway = hp["min_highway_ratio"], hp["max_highway_ratio"]
    cfg.range_lo, cfg.range_hi = hp["range"]
    cfg.clip_value = hp["clip_grad_value"] if hp["clip_grads"] else 0.0
    cfg.adam_beta1, cfg.adam_beta2, cfg.adam_eps = 0.9, 0.999, 1e-8      # tf.train.AdamOptimizer defaults (:1267)
    cfg.max_batch = int(max_batch)
    cfg.train_capacity = int(bool(train))
    cfg.operand_dtype = {"fp32": _cabi.OPERAND_FP32, "bf16": _cabi.OPERAND_BF16}[operand_dtype]
    cfg.share_theta_weights = int(bool(hp.get("share_theta_weights", False)))
    cfg.share_phi_weights = int(bool(hp.get("share_phi_weights", False)))
    cfg.add_noise_to_chain = int(bool(hp.get("add_noise_to_chain", False)))
    if cfg.add_noise_to_chain:
        for t, v in enumerate(hp["noise_stddevs"]):
            cfg.noise_stddevs[t] = float(v)
    return cfg
