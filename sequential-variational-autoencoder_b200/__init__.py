"""seqvae_b200 - B200-native Sequential-VAE hot path behind the reference's SequentialVAE surface.

Host side (this package, Python like the reference) -> thin ctypes binding (``_cabi``) -> ``libsvae.so`` (C ABI,
``include/svae.h``) -> hand-written sm_100a CUDA kernels (``csrc/``).  There is no CPU fallback: constructing a model
without the built library or without a CUDA device raises.
"""
from .config import hyperparams, NETNAMES  # noqa: F401
from .dataset import Dataset, SyntheticDataset  # noqa: F401
from .sequential_vae import SequentialVAE  # noqa: F401
from .trainer import NoisyTrainer  # noqa: F401
from . import _cabi  # noqa: F401

__all__ = ["SequentialVAE", "NoisyTrainer", "SyntheticDataset", "Dataset", "hyperparams", "NETNAMES"]
