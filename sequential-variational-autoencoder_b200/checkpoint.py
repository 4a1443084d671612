"""Checkpoint layout: the variable set ``tf.train.Saver()`` writes for this graph (reference abstract_network.py:124-152),
so that weights are exchangeable name-for-name with anyone who can run the TensorFlow 1.x reference.

``tf.train.Saver()`` with no arguments saves every global variable:
  * the trainable variables (SURVEY App. D names; a homogeneous chain lists its shared scopes once);
  * per batch-norm layer the non-trainable ``moving_mean`` / ``moving_variance`` - created by
    ``tf.contrib.layers.batch_norm`` but never updated and never read by this model (``is_training`` stays
    True and the train op does not run UPDATE_OPS), so they keep their initial values zeros / ones;
  * the Adam slots ``<var>/Adam`` (m) and ``<var>/Adam_1`` (v) of every variable that has a gradient - the dead
    recognition branch gets ``None`` gradients, ``apply_gradients`` skips it and creates no slots;
  * ``beta1_power`` / ``beta2_power``, which hold beta^(t+1) after t updates (initialised to beta, multiplied once per step).
Host-side schedule state the reference forgets to save (``iteration``, ``learning_rate``: sequential_vae.py:1351-1357
restarts the KL warm-up on every resume) is stored under ``__``-prefixed keys that a TF reader would ignore.
"""
import math

PF_INERT, PF_DEAD = 2, 4


def tf_checkpoint_layout(param_table, train=True):
    """[(name, shape, kind, param_name)] in save order.  kind: "param" | "bn_moving_mean" | "bn_moving_variance" |
    "adam_m" | "adam_v" | "beta1_power" | "beta2_power".  ``param_table``: SequentialVAE.param_table (or the device-free
    svae_param_table) as dicts with name / shape / flags."""
    out = []
    for p in param_table:
        out.append((p["name"], tuple(p["shape"]), "param", p["name"]))
        if p["name"].endswith("/beta"):
            scope = p["name"][: -len("/beta")]
            out.append((scope + "/moving_mean", tuple(p["shape"]), "bn_moving_mean", p["name"]))
            out.append((scope + "/moving_variance", tuple(p["shape"]), "bn_moving_variance", p["name"]))
    if train:
        out.append(("beta1_power", (), "beta1_power", None))
        out.append(("beta2_power", (), "beta2_power", None))
        for p in param_table:
            if p["flags"] & PF_DEAD:
                continue
            out.append((p["name"] + "/Adam", tuple(p["shape"]), "adam_m", p["name"]))
            out.append((p["name"] + "/Adam_1", tuple(p["shape"]), "adam_v", p["name"]))
    return out


def beta_powers(adam_t, beta1=0.9, beta2=0.999):
    """Values of the beta1_power / beta2_power variables after ``adam_t`` updates."""
    return beta1 ** (adam_t + 1), beta2 ** (adam_t + 1)


def adam_t_from_beta_powers(beta1_power, beta2_power, beta1=0.9, beta2=0.999):
    """Inverse of ``beta_powers`` for a checkpoint written by TensorFlow (no ``__adam_t`` key).  The fp32 ``beta1_power``
    = 0.9^(t+1) underflows to 0 after ~1000 updates, i.e. for any real checkpoint; ``beta2_power`` = 0.999^(t+1) resolves the
    step count up to ~88 000 updates in fp32 and is used whenever it is positive.  Beyond that both bias corrections equal 1
    to fp32 precision, so any large count gives the same update: 10^6 is returned."""
    for power, beta in ((beta2_power, beta2), (beta1_power, beta1)):
        if power is None:
            continue
        p = float(power)
        if p > 0.0 and p < 1.0:
            return max(0, int(round(math.log(p) / math.log(beta))) - 1)
        if p >= 1.0:
            return 0
    return 1000000
