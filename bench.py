#!/usr/bin/env python
"""bench.py - headline benchmark of the Sequential-VAE hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      (CPU arm: the oracle port of the reference graph)

Metric (BASELINE.json): training images/s (forward + per-step ELBO + full backward + clipped Adam) of the CelebA-64
default SequentialVAE, batch 100 per GPU (weak scaling, per-replica batch-norm), synthetic data, random-init weights.
A "step" is one pass of the hot path over one batch.  ``value`` is measured with the batch resident in HBM;
``e2e`` goes through the reference-shaped public API (``SequentialVAE.train(numpy, numpy)``) with host buffers, i.e.
host->device copies of the inputs and the device->host read of the losses inside the timed region.

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (netname, dataset shape name, batch per GPU, hyper-parameter overrides)
    "celeba64_b100": ("c_inhomog", "celebA", 100, {}),
    "mnist32_b100": ("m_inhomog", "mnist", 100, {}),
    "cifar32_b100": ("c_inhomog", "cifar", 100, {}),
    "lsun64_b256_t16": ("sequential_vae_lsun", "lsun", 256, {"mc_steps": 16}),
    "lsun64_b256_t25": ("sequential_vae_lsun", "lsun", 256, {"mc_steps": 25}),      # BASELINE configs[3]: "long chain"
    # homogeneous (weight-shared) chains, SURVEY 8 f1: the same per-image work as celeba64_b100 at T=8
    "celeba64_b100_homog": ("sequential_vae_celebA_homog", "celebA", 100, {}),
    "celeba64_b100_homog_t25": ("c_homog", "celebA", 100, {}),
}
# Algorithmic work per image (DESIGN.md "Roofline arithmetic")
ALGO = {
    "celeba64_b100": dict(train_flops=13.0065e9, train_bytes=62.88e6, gen_flops=3.2303e9, gen_bytes=9.54e6),
    "mnist32_b100": dict(train_flops=4.9002e9, train_bytes=16.46e6),
    "cifar32_b100": dict(train_flops=3.2517e9, train_bytes=22.56e6),
    "lsun64_b256_t16": dict(train_flops=26.8676e9, train_bytes=101.42e6),
    # the T = 16 figures scaled by 25/16 (step 0 has no chain encoder: overestimates the work by < 0.5 %)
    "lsun64_b256_t25": dict(train_flops=26.8676e9 * 25 / 16, train_bytes=101.42e6 * 25 / 16),
    # 5 x 3.294 M activation elements x 2 B + 40 B x 14.87 M live (shared) parameters / 100
    "celeba64_b100_homog": dict(train_flops=13.0065e9, train_bytes=38.89e6),
}
METRIC = "train img/s (fwd+bwd+ELBO) CelebA-64 SeqVAE @1/2/4/8 B200; sample img/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")   # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm),
                    power_w_max=max(power))


def cpu_oracle_arm(workload, steps, warmup, threads=None, budget_s=None):
    """The reference's CPU path: fp32 PyTorch-CPU restatement of the reference graph (TF 1.x is not installable,
    SURVEY Q14), one full train step = forward over all T steps + autograd backward + clip + TF-Adam.  budget_s: stop
    timing new steps once this much wall time has been spent (at least one timed step always runs); the number of steps
    actually timed is returned."""
    import numpy as np
    import torch

    from oracle import seqvae_oracle as O
    from seqvae_b200.dataset import _SHAPES

    netname, shape, B, over = WORKLOADS[workload]
    dims, rng = _SHAPES[shape]
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = O.hyperparams(netname, dims, rng, **over)
    om = O.OracleModel(hp, seed=0, dtype=torch.float32)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand([B] + dims, generator=g) * (rng[1] - rng[0]) + rng[0]
    times = []
    t_begin = time.perf_counter()
    warm_done = 0
    for i in range(warmup + steps):
        eps = torch.randn(hp["mc_steps"], B, hp["latent_dim"], generator=g)
        t0 = time.perf_counter()
        om.train(x, x, eps)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
        else:
            warm_done += 1
        if budget_s is not None and times and time.perf_counter() - t_begin > budget_s:
            break
    total = sum(times)
    mean = total / len(times)
    return dict(value=B / mean, ms_per_step=mean * 1e3, cores=cores, B=B, steps=len(times), warmup=warm_done,
                mc_steps=hp["mc_steps"], latent_dim=hp["latent_dim"], image=list(dims),
                sample="%d full train steps (B=%d, T=%d, fp32, forward + autograd backward + clip + TF-Adam) after %d warm-up"
                       % (len(times), B, hp["mc_steps"], warm_done))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, on this box's host cores.  The reference
    itself needs TensorFlow 1.x (tf.contrib) which cannot be installed for Python 3.12, so this is the oracle PORT.
    Honours --steps / --warmup (a wall-time budget of --ref-budget seconds only cuts the run short on a very slow host);
    under torchrun (N > 1) rank 0 alone runs ONE CPU process on its host cores - the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_oracle_arm(args.workload, args.steps, args.warmup, budget_s=args.ref_budget)
    netname, shape, B, over = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "img/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "netname": netname, "image": r["image"], "batch_per_gpu": B,
                   "global_batch": B, "mc_steps": r["mc_steps"], "latent_dim": r["latent_dim"],
                   "note": "CPU oracle port of the reference graph (TensorFlow 1.x not installable); ONE CPU process on rank "
                           "0's host cores whatever --gpus is: at N > 1 compare it with the per-GPU rate, not the aggregate"},
        "cpu_baseline": {"value": r["value"], "unit": "img/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    # stdout carries exactly ONE JSON line: libraries (NCCL's version banner, ...) that write to fd 1 go to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="celeba64_b100", choices=sorted(WORKLOADS))
    ap.add_argument("--operand", default=os.environ.get("SVAE_OPERAND", "auto"), choices=["auto", "fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-generation", action="store_true")
    ap.add_argument("--gen-batch", type=int, default=4096)
    ap.add_argument("--profile-json", default=None, help="also write the per-kernel-class table to this file")
    ap.add_argument("--ref-budget", type=float, default=300.0, help="--impl reference: wall-time cap in seconds")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import seqvae_b200 as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    netname, shape, B, over = WORKLOADS[args.workload]
    operand = args.operand
    if operand == "auto":
        operand = "bf16"
    ds = S.SyntheticDataset(shape, B, seed=1234 + rank)
    model = S.SequentialVAE(ds, B, netname, device=local_rank, operand_dtype=operand, restore=False, seed=0, **over)
    if model.tc_layers == 0:
        operand_used = "fp32"      # no contraction was routed to the tensor-core kernels
    else:
        operand_used = operand
    stream = torch.cuda.Stream(device=local_rank, priority=-1)   # the chain runs on it: highest priority, side streams lowest
    model.use_torch_stream(stream)
    if world > 1:
        from seqvae_b200.dist import attach_communicator

        attach_communicator(model, dist, rank, world)

    H, W, C = model.data_dims
    T, Z = model.mc_steps, model.latent_dim
    x_host = ds.next_batch(B)
    tgt_host = x_host.copy()
    with torch.cuda.stream(stream):
        x_dev = torch.from_numpy(x_host).cuda(local_rank, non_blocking=False)
        tgt_dev = x_dev.clone()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ------------------------------------------------------------------------
    for _ in range(args.warmup):
        model.train_async(x_dev, tgt_dev)          # eps: in-kernel Philox (benchmark mode)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = model.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        model.train_async(x_dev, tgt_dev)          # CUDA-graph replay of the forked (4-stream) step
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = model.launch_count - l0
    # Per-kernel durations: the same steps once more with every launch bracketed by CUDA events.  Profiling serialises
    # the step on one stream without the graph (events cannot be timed inside a captured graph and concurrent kernels
    # would inflate each other's durations), so it is kept OUT of the timed region above.
    prof_steps = max(1, min(args.steps, 5))
    model.profile(True)
    for _ in range(prof_steps):
        model.train_async(x_dev, tgt_dev)
    prof = model.profile_read()
    model.profile(False)
    barrier()
    clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end to end through the public API with host buffers ("e2e") ---------------------------------------------------
    e2e_steps = max(3, min(args.steps, 10))

    def time_e2e(xh, th):
        for _ in range(2):
            model.train(xh, th)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            model.train(xh, th)                   # numpy in, float out: H2D of x and target, D2H of the losses, sync
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # (a) ordinary pageable numpy arrays - what trainer.py:100-104 passes (dataset.next_batch output): the library stages
    #     them through its pinned buffer; (b) page-locked arrays (numpy views of torch pinned tensors): DMA straight from them
    e2e_pageable_ms = time_e2e(x_host, tgt_host)
    x_pin = torch.from_numpy(x_host).pin_memory().numpy()
    tgt_pin = torch.from_numpy(tgt_host).pin_memory().numpy()
    e2e_ms = time_e2e(x_pin, tgt_pin)
    e2e = {"value": world * B / (e2e_ms * 1e-3), "unit": "img/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(2 * x_host.nbytes), "d2h_bytes_per_step": int(4 * (2 + 2 * 64)),
           "host_buffers": "pinned (page-locked) numpy arrays, copied host->device inside every timed call",
           "pageable": {"value": world * B / (e2e_pageable_ms * 1e-3), "ms_per_step": e2e_pageable_ms,
                        "host_buffers": "ordinary pageable numpy arrays (what trainer.py:100-104 passes), staged through "
                                        "the library's pinned buffer inside every timed call"}}

    # ---- data-parallel consistency: after all those updates every rank must hold bit-identical parameters ---------------
    dp_check = None
    if dist is not None:
        import zlib

        arena = model.read_arena("param")
        mine = (int(zlib.crc32(arena.tobytes())), float(arena.astype(np.float64).sum()))
        allv = [None] * world
        dist.all_gather_object(allv, mine)
        dp_check = {"ranks": world, "param_crc32": [v[0] for v in allv], "ranks_agree": len({v[0] for v in allv}) == 1,
                    "train_steps_before_check": int(model.iteration), "param_sum": allv[0][1]}
        del arena

    # ---- roofline of the dominant kernel class (live CUDA-event durations from the timed region) -------------------------
    peaks = measured_peaks()
    roof = None
    table = []
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        table.append(dict(kernel=name, launches_per_step=v["launches"] / prof_steps, ms_per_step=v["ms"] / prof_steps,
                          share=v["ms"] / tot_ms, tflops=v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] else 0.0,
                          gbs=v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] else 0.0))
    if table:
        top = table[0]
        contraction = top["kernel"].startswith(("gather_gemm", "wgrad"))
        if contraction:
            roof = {"kernel": top["kernel"], "bound": "tensor", "achieved": top["tflops"], "peak": peaks["tf_sustained"],
                    "unit": "TFLOP/s", "frac": top["tflops"] / peaks["tf_sustained"], "traffic": None,
                    "share_of_step": top["share"], "peak_source": peaks["source"] + " bf16 sustained"}
        else:
            roof = {"kernel": top["kernel"], "bound": "hbm", "achieved": top["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": top["gbs"] / peaks["hbm"], "traffic": None, "share_of_step": top["share"],
                    "peak_source": peaks["source"]}
    # DRAM traffic per launch of that kernel class, from the committed `ncu --set full` capture (profiles/traffic.json is
    # written by scripts/ncu_raw_summary.py from the .ncu-rep of the same bench command)
    if roof is not None:
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get(roof["kernel"])
            if ent:
                roof["traffic"] = ent["dram_bytes_per_launch"]
                roof["traffic_source"] = ent.get("source")
        except (OSError, ValueError):
            pass
    step_traffic = None
    try:      # whole-step DRAM bytes from the committed ncu launch list of the same command (profiles/rN_launches.md)
        step_traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("step")
    except (OSError, ValueError):
        pass
    parity = None
    try:      # measured by tests/test_gpu_fullsize.py::test_full_size_oracle_parity on a B200 (committed copy under profiles/)
        parity = json.load(open(os.path.join(ROOT, "profiles", "parity_fullsize.json")))
        parity["source"] = "profiles/parity_fullsize.json (tests/test_gpu_fullsize.py::test_full_size_oracle_parity, B200)"
    except (OSError, ValueError):
        pass
    algo = ALGO.get(args.workload, {})
    path_roof = None
    if algo:
        per_gpu = value / world
        path_roof = {"tensor_frac": per_gpu * algo["train_flops"] / (peaks["tf_sustained"] * 1e12),
                     "hbm_frac": per_gpu * algo["train_bytes"] / (peaks["hbm"] * 1e9),
                     "flops_per_img": algo["train_flops"], "bytes_per_img": algo["train_bytes"]}

    # ---- generation-mode chain (config 5: replicas only) -----------------------------------------------------------------
    generation = None
    if not args.no_generation and args.workload == "celeba64_b100":
        model.close()
        del model
        torch.cuda.empty_cache()
        GB = args.gen_batch
        gmodel = S.SequentialVAE(ds, GB, netname, device=local_rank, operand_dtype=operand, restore=False, seed=0,
                                 train=False, **over)
        gmodel.use_torch_stream(stream)
        with torch.cuda.stream(stream):
            out = torch.empty([T, GB, H, W, C], device="cuda", dtype=torch.float32)
        for i in range(2):
            gmodel.generate_async(GB, out, None, seed=i)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gsteps = 3
        gl0 = gmodel.launch_count
        g0.record(stream)
        for i in range(gsteps):
            gmodel.generate_async(GB, out, None, seed=10 + i)
        g1.record(stream)
        barrier()
        gms = g0.elapsed_time(g1) / gsteps
        t = torch.tensor([gms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gms = float(t.item())
        gval = world * GB / (gms * 1e-3)
        generation = {"metric": "sample img/s (generation-mode chain, all %d steps)" % T, "value": gval, "unit": "img/s",
                      "batch_per_gpu": GB, "ms_per_chain": gms, "scaling": "replicas only",
                      "gpu_launches": gmodel.launch_count - gl0,
                      "tensor_frac": gval / world * algo["gen_flops"] / (peaks["tf_sustained"] * 1e12),
                      "hbm_frac": gval / world * algo["gen_bytes"] / (peaks["hbm"] * 1e9)}
        gmodel.close()

    # ---- CPU baseline (rank 0, N=1 only; bounded sample) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_oracle_arm(args.workload, 2, 1)
        cpu = {"value": r["value"], "unit": "img/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "ms_per_step": r["ms_per_step"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if operand_used == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": args.workload, "netname": netname, "image": [H, W, C], "batch_per_gpu": B,
                       "global_batch": B * world, "mc_steps": T, "latent_dim": Z,
                       "parallelism": "dp%d (per-replica BN, NCCL grad all-reduce per chain step)" % world,
                       "operands": "bf16 tcgen05 + fp32 accumulate" if operand_used == "bf16" else "fp32 SIMT",
                       "l2": "per-step working set (activations+weights+Adam state, >3 GB) exceeds the 126 MB L2",
                       "eps": "in-kernel Philox",
                       "execution": "CUDA graph of the whole step (chain stream with programmatic dependent launch; side "
                                    "streams: the chain's weight gradients, the latent projections, the recognition nets of "
                                    "all chain steps as batched launches (blockIdx.z = chain step) in groups, their weight "
                                    "gradients, per-chain-step Adam + operand repack); `kernels`/`roofline` come from %d "
                                    "extra serialised, event-bracketed steps right after the timed region" % prof_steps},
            "roofline": roof, "path_roofline": path_roof, "kernels": table, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "generation": generation,
            "step_traffic": step_traffic if args.workload == "celeba64_b100" else None,
            "parity": parity if args.workload == "celeba64_b100" else None, "dp_consistency": dp_check,
        }
        print(json.dumps(line), flush=True)
        if args.profile_json:
            with open(args.profile_json, "w") as f:
                json.dump(line, f, indent=1)
    try:
        model.close()                 # destroys the captured graphs and the library's communicator before torch's
    except NameError:
        pass
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
