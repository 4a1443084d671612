"""Worker for the data-parallel tests (launched by torch.distributed.run, one process per rank).

backend nccl (GPU): each rank trains ONE step on its shard through libsvae with the in-library NCCL all-reduce; rank 0
also recomputes both shards' gradients on a communicator-free handle, averages them on the host, applies clip + TF-Adam
in numpy and compares every live parameter.
backend gloo (CPU): the same protocol on the oracle (the product has no CPU path): per-replica batch-norm, gradients
averaged over ranks with all_reduce, identical weights afterwards - this pins the SEMANTICS the GPU path implements and
exercises shard_batch / the rendezvous plumbing without a GPU.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import seqvae_oracle as O  # noqa: E402
from oracle import tf_semantics_np as TFNP  # noqa: E402
import seqvae_b200 as S  # noqa: E402
from seqvae_b200.dist import attach_communicator, shard_batch  # noqa: E402

OVER = dict(filter_sizes=[3, 8, 16, 16, 24, 24], vlae_latent_dims=[2, 3, 2, 2], mc_steps=2)
DIMS, RNG, GB = [16, 16, 3], (-1.0, 1.0), 12
OPERAND = sys.argv[4] if len(sys.argv) > 4 else "fp32"
if OPERAND == "bf16":   # every conv on the TMA-fed tcgen05 kernels: the recognition nets run as batched launches and the
    # gradient slice of a chain step is all-reduced in two halves (chain encoder + decoder, recognition net)
    OVER = dict(filter_sizes=[3, 32, 32, 64, 64, 64], vlae_latent_dims=[2, 3, 2, 2], mc_steps=3)
    DIMS = [32, 32, 3]


def inputs(hp):
    g = torch.Generator().manual_seed(5)
    x = (torch.rand([GB] + DIMS, generator=g, dtype=torch.float64) * 2 - 1).float().double()
    eps = torch.randn(hp["mc_steps"], GB, hp["latent_dim"], generator=g, dtype=torch.float64).float().double()
    return x, eps


def main():
    import faulthandler

    faulthandler.dump_traceback_later(150, exit=True)   # a hung collective must become a failed test, not a stuck box
    backend = sys.argv[1]
    out_path = sys.argv[2]
    # optional third argument: netname (a homogeneous chain exercises the tied slices under the all-reduce)
    net = sys.argv[3] if len(sys.argv) > 3 else "c_inhomog"
    if net != "c_inhomog":
        OVER["mc_steps"] = 3
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group(backend)
    hp = O.hyperparams(net, DIMS, RNG, **OVER)
    P = O.init_params(hp, 0)
    x, eps = inputs(hp)
    lo, hi = shard_batch(GB, rank, world)
    xs, es = x[lo:hi], eps[:, lo:hi]
    result = {}
    if backend == "gloo":
        fw, grads = O.loss_and_grads(hp, P, xs, xs, es, 1.0)
        avg = {}
        for k, g in grads.items():
            if g is None:
                avg[k] = None
                continue
            t = g.clone()
            dist.all_reduce(t)
            avg[k] = t / world
        st = O.AdamState(P)
        with torch.no_grad():
            O.adam_apply(hp, P, avg, st, 2e-4, update_inert=False)
        # every rank must hold identical weights; rank 0 checks against a single-process emulation of all shards
        flat = torch.cat([v.reshape(-1) for v in P.values()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(flat, ref))
        if rank == 0:
            P2 = O.init_params(hp, 0)
            gs = []
            for r in range(world):
                a, b = shard_batch(GB, r, world)
                gs.append(O.loss_and_grads(hp, P2, x[a:b], x[a:b], eps[:, a:b], 1.0)[1])
            from seqvae_b200.dist import average_gradients_reference
            avg2 = average_gradients_reference(gs)
            st2 = O.AdamState(P2)
            with torch.no_grad():
                O.adam_apply(hp, P2, avg2, st2, 2e-4, update_inert=False)
            err = max(float((P[k] - P2[k]).abs().max()) for k in P)
            result = dict(same=same, err=err, shards=[shard_batch(GB, r, world) for r in range(world)])
        allsame = torch.tensor([1 if same else 0])
        dist.all_reduce(allsame, op=dist.ReduceOp.MIN)
        if rank == 0:
            result["all_same"] = bool(allsame.item())
    else:
        dev = int(os.environ["LOCAL_RANK"])
        ds = S.SyntheticDataset("x", hi - lo, data_dims=DIMS, data_range=list(RNG))
        model = S.SequentialVAE(ds, hi - lo, net, device=dev, operand_dtype=OPERAND, restore=False, **OVER)
        model.set_params({k: v.numpy() for k, v in P.items()})
        attach_communicator(model, dist, rank, world)
        model.iteration = 4999          # reg_coeff = 1 - exp(-1) on the step below
        model.train(xs.numpy().astype(np.float32), xs.numpy().astype(np.float32), es.numpy())
        got = model.get_params(live_only=True)
        flat = torch.from_numpy(np.concatenate([v.reshape(-1) for v in got.values()])).cuda()
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([1 if torch.equal(flat, ref) else 0], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if rank == 0:
            # emulate on one communicator-free handle: gradients of every shard, host average, numpy clip + TF-Adam
            reg = 1 - np.exp(-1.0)
            gsum = None
            for r in range(world):
                a, b = shard_batch(GB, r, world)
                dsr = S.SyntheticDataset("x", b - a, data_dims=DIMS, data_range=list(RNG))
                m2 = S.SequentialVAE(dsr, b - a, net, device=dev, operand_dtype=OPERAND, restore=False, **OVER)
                m2.set_params({k: v.numpy() for k, v in P.items()})
                m2.forward(x[a:b].numpy(), None, eps[:, a:b].numpy(), reg)
                m2.backward()
                g = m2.gradients(live_only=True)
                gsum = g if gsum is None else {k: gsum[k] + g[k] for k in g}
                m2.close()
            err, errs = 0.0, []
            for k, g in gsum.items():
                gk = np.clip(g.astype(np.float64) / world, -10, 10)
                p1, _, _ = TFNP.adam_tf(P[k].numpy(), gk, np.zeros_like(gk), np.zeros_like(gk), 1, 2e-4)
                upd_ref = p1 - P[k].numpy()
                upd = got[k].astype(np.float64) - P[k].numpy()
                if np.linalg.norm(upd_ref) > 1e-9:
                    errs.append(float(np.linalg.norm(upd - upd_ref) / np.linalg.norm(upd_ref)))
            # Adam's first update is ~lr * sign(g): an element whose gradient is at the rounding noise may move the other way, which
            # is O(1) on a small tensor - the 90th percentile over the tensors is the robust statement, the maximum only a sanity bound
            err = float(np.percentile(errs, 90)) if errs else 0.0
            result = dict(all_same=bool(same.item()), err=err, err_max=max(errs) if errs else 0.0, loss=model.last_losses["loss"])
        # three more steps: from the second step on the library replays a captured CUDA graph that contains the NCCL
        # all-reduces; every rank must still hold bit-identical weights, and they must have moved
        for _ in range(3):
            model.train(xs.numpy().astype(np.float32), xs.numpy().astype(np.float32), es.numpy())
        got2 = model.get_params(live_only=True)
        flat2 = torch.from_numpy(np.concatenate([v.reshape(-1) for v in got2.values()])).cuda()
        ref2 = flat2.clone()
        dist.broadcast(ref2, 0)
        same2 = torch.tensor([1 if (torch.equal(flat2, ref2) and not torch.equal(flat2, flat) and bool(torch.isfinite(flat2).all())) else 0],
                             device="cuda")
        dist.all_reduce(same2, op=dist.ReduceOp.MIN)
        if rank == 0:
            result["all_same_after_graph_steps"] = bool(same2.item())
        # homogeneous chain: the per-step slices of every shared variable must still be bit-identical on every rank
        arena = model.read_arena("param")
        tied = 1
        for p in model.param_table:
            offs = model.param_slices(p["name"])
            for o in offs[1:]:
                if not np.array_equal(arena[offs[0]:offs[0] + p["numel"]], arena[o:o + p["numel"]]):
                    tied = 0
        tied = torch.tensor([tied], device="cuda")
        dist.all_reduce(tied, op=dist.ReduceOp.MIN)
        if rank == 0:
            result["slices_tied"] = bool(tied.item())
            result["shared_variables"] = sum(1 for p in model.param_table if len(model.param_slices(p["name"])) > 1)
        model.close()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(result, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
