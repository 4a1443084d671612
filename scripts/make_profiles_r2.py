"""Turn the outputs of scripts/gpu_evidence_r2.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/.
usage: python scripts/make_profiles_r2.py r2a r2"""
import collections, csv, json, sys
src, tag = sys.argv[1], sys.argv[2]
G = "gpurun_out/%s_" % src

line = open(G + "bench_final.json").read().strip().splitlines()[-1]
open("profiles/%s_bench_final.json" % tag, "w").write(line + "\n")
open("profiles/%s_bench_reference_arm.json" % tag, "w").write(open(G + "bench_reference.json").read().strip().splitlines()[-1] + "\n")
d = json.loads(line)


def scale(v, u):
    return v * {"us": 1e3, "usecond": 1e3, "ms": 1e6, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u.split("/")[0], 1.0)


# ---- launch list with DRAM bytes
rows = list(csv.DictReader([l for l in open(G + "launches.csv") if l.startswith('"')]))
per = collections.OrderedDict()
for r in rows:
    p = per.setdefault(int(r["ID"]), {"k": r["Kernel Name"], "g": r["Grid Size"]})
    p[r["Metric Name"]] = scale(float(r["Metric Value"]), r["Metric Unit"])


def short(n):
    n = n.replace("<unnamed>::", "").replace("void ", "")
    if "svae_multi_kernel" in n:
        for b in ("bn_act_fwd_v8", "bn_act_fwd", "bn_bwd_reduce_v8", "bn_bwd_reduce", "bn_bwd_apply_v8", "bn_bwd_apply", "reparam_fwd", "reparam_bwd",
                  "heads_fwd_tiled", "heads_fwd", "heads_dgrad_tiled", "heads_dgrad", "heads_wgrad_tiled", "heads_wgrad", "lat_fwd_moment", "lat_fwd_fused", "lat_bwd_fused"):
            if b + "_kernel_body" in n:
                return "multi<" + b + ">"
        return "multi<?>"
    return n.split("(")[0][:48]


agg = collections.OrderedDict()
for p in per.values():
    a = agg.setdefault(short(p["k"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += p["gpu__time_duration.sum"]; a[2] += p.get("dram__bytes_read.sum", 0.0); a[3] += p.get("dram__bytes_write.sum", 0.0)
T = sum(a[1] for a in agg.values()); R = sum(a[2] for a in agg.values()); W = sum(a[3] for a in agg.values())
md = """# Round 2: ncu launch list of ONE graph-replayed training step with per-launch DRAM traffic (CelebA-64, B=100, T=8, bf16 tcgen05)

Command (scripts/gpu_evidence_r2.sh): `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
--cache-control none -s 2550 -c 842 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation`, right after the same
command exited 0 without ncu.  842 launches = one step (round 1: 1 167).  Per-launch times under ncu are serialised: compare SHARES
with the `kernels` table of `profiles/%s_bench_final.json` (same build, no profiler: %.2f ms/step, %.0f img/s).

Whole step: %d launches, %.2f ms of kernel time, **DRAM read %.2f GB + write %.2f GB = %.2f GB** against 6.29 GB algorithmic
(SURVEY App. F: bf16 activations + 40 B per live parameter): %.1fx.  The excess is the fp32 pre-batch-norm tensors and fp32 gradient
tensors (the algorithmic figure assumes bf16 storage), the partial-sum flushes of the weight gradients and the weight re-reads of the
small-map layers; under ncu every kernel also starts with whatever the serialised predecessor left in L2, not with the overlapped
step's residency.

| kernel | launches | time us | share %% | DRAM read MB | DRAM write MB |
|---|---|---|---|---|---|
""" % (tag, d["ms_per_step"], d["value"], len(per), T / 1e6, R / 1e9, W / 1e9, (R + W) / 1e9, (R + W) / 6.29e9)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md += "| %s | %d | %.1f | %.1f | %.1f | %.1f |\n" % (k, a[0], a[1] / 1e3, 100 * a[1] / T, a[2] / 1e6, a[3] / 1e6)
open("profiles/%s_launches.md" % tag, "w").write(md)

# ---- full captures
WANT = [("grid", "Grid Size"), ("time us", "gpu__time_duration.sum"), ("DRAM rd MB", "dram__bytes_read.sum"), ("DRAM wr MB", "dram__bytes_write.sum"),
        ("tensor pipe % (active)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("tensor pipe % (elapsed)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
        ("dyn smem KB", "launch__shared_mem_per_block_dynamic"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
        ("SM active cycles", "sm__cycles_active.avg"), ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed")]
TITLES = {"wgrad": ("`tc2_wgrad_kernel` (TMA-fed tcgen05 weight gradient, tap-stacked, cut chosen by the cost model)", "-k regex:tc2_wgrad_kernel -s 60 -c 10"),
          "bn": ("the three batch-norm kernels (`bn_act_fwd_v8`, `bn_bwd_reduce_v8`, `bn_bwd_apply_v8`)", "-k regex:'bn_bwd_reduce_v8_kernel|bn_bwd_apply_v8_kernel|bn_act_fwd_v8_kernel' -s 300 -c 12"),
          "conv": ("`tc2_conv_kernel` (TMA-fed tcgen05 implicit GEMM: conv, deconv, input gradients)", "-k regex:tc2_conv_kernel -s 150 -c 12"),
          "multi": ("the batched (blockIdx.z = chain step) launches of the recognition nets", "-k regex:multi_kernel -s 60 -c 14")}
traffic = {}
for name, (title, sel) in TITLES.items():
    rows = list(csv.reader(open(G + "raw_%s.csv" % name)))
    h, u = rows[0], rows[1]
    md = "# Round 2: `ncu --set full` capture of %s\n\nCommand: `ncu --set full --clock-control none --import-source on %s python bench.py --steps 2 --warmup 3\n--no-cpu-baseline --no-generation` (CelebA-64, B=100, T=8); raw page exported on the GPU box with `ncu -i ... --page raw --csv`\n(scripts/gpu_evidence_r2.sh; the .ncu-rep files exceed what the box copies back), reduced by scripts/make_profiles_r2.py.\n\n" % (title, sel)
    md += "| kernel | " + " | ".join(w[0] for w in WANT) + " |\n|" + "---|" * (len(WANT) + 1) + "\n"
    tot = 0.0; n = 0
    for r in rows[2:]:
        vals = []
        for lab, key in WANT:
            i = h.index(key)
            v = r[i]
            if key.startswith("dram__bytes"):
                b = scale(float(v), u[i]); vals.append("%.2f" % (b / 1e6)); tot += b
            elif key == "launch__shared_mem_per_block_dynamic":
                vals.append("%.0f" % (scale(float(v), u[i]) / 1e3))
            elif key == "gpu__time_duration.sum":
                vals.append("%.1f" % (scale(float(v), u[i]) / 1e3))
            elif key in ("Grid Size",):
                vals.append(v.replace(" ", ""))
            else:
                vals.append("%.1f" % float(v) if v.replace(".", "").isdigit() else v)
        n += 1
        md += "| %s | " % short(r[h.index("Kernel Name")]) + " | ".join(vals) + " |\n"
    md += "\nMean DRAM traffic per launch: %.2f MB over %d launches.\n" % (tot / max(n, 1) / 1e6, n)
    traffic[name] = tot / max(n, 1)
    open("profiles/%s_ncu_%s.md" % (tag, name), "w").write(md)
def nlaunch(name):
    return len(list(csv.reader(open(G + "raw_%s.csv" % name)))) - 2


json.dump({"gather_gemm_tcgen05": {"dram_bytes_per_launch": traffic["conv"], "launches_captured": nlaunch("conv"),
                                   "source": "profiles/%s_ncu_conv.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches of tc2_conv_kernel)" % tag},
           "wgrad_tcgen05": {"dram_bytes_per_launch": traffic["wgrad"], "launches_captured": nlaunch("wgrad"),
                             "source": "profiles/%s_ncu_wgrad.md (same metrics, tc2_wgrad_kernel)" % tag},
           "step": {"dram_bytes": R + W, "dram_bytes_read": R, "dram_bytes_write": W, "launches": len(per), "algorithmic_bytes": 6.288e9,
                    "source": "profiles/%s_launches.md (ncu launch list of one graph-replayed step, --cache-control none)" % tag}},
          open("profiles/traffic.json", "w"), indent=1)
print("ok", traffic)
