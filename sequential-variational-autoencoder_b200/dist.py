"""Data-parallel plumbing: one process per GPU (torch.distributed for rendezvous / barriers), gradients all-reduced by
NCCL *inside* libsvae (one bucket per chain step, issued on a side stream as soon as that step's backward ends).

The reference has no multi-device path for this model; the semantics defined here: batch-sharded
replicas with per-replica batch-norm statistics (each replica is numerically the single-GPU reference step on its
shard), gradients averaged over ranks before clipping and Adam, identical initial weights on every rank."""
import ctypes as C

from . import _cabi


def shard_batch(global_batch, rank, world):
    """Rows [lo, hi) of a global batch owned by `rank` (contiguous, remainder spread over the first ranks)."""
    base, rem = divmod(int(global_batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def attach_communicator(model, dist, rank, world):
    """Create the NCCL communicator of `model`'s handle.  `dist` is an initialised torch.distributed module (any
    backend): it only carries the 128-byte ncclUniqueId from rank 0 to the others."""
    L = _cabi.lib()
    path = _cabi.nccl_library_path()
    cpath = path.encode() if path else None
    ident = [None]
    if rank == 0:
        buf = C.create_string_buffer(128)
        _cabi.check(None, L.svae_nccl_unique_id(buf, cpath))
        ident = [buf.raw]
    dist.broadcast_object_list(ident, src=0)
    _cabi.check(model._h, L.svae_comm_init(model._h, int(rank), int(world), ident[0], cpath))
    return model


def average_gradients_reference(grad_dicts):
    """What the in-library all-reduce computes, restated on host arrays (used by the gloo CPU tests): the mean over
    ranks of every gradient tensor."""
    n = len(grad_dicts)
    out = {}
    for k in grad_dicts[0]:
        vals = [g[k] for g in grad_dicts]
        out[k] = None if vals[0] is None else sum(vals) / n
    return out
