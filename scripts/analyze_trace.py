"""Analyse gpurun_out/trace_step.json (scripts/trace_step.py): per-stream busy time and the biggest idle gaps of the chain."""
import collections, json, sys
ev = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_step.json"))
# split the two steps at the largest-work boundary: take the second half by time
span = ev[-1]["t"] + ev[-1]["d"]
streams = collections.defaultdict(list)
for e in ev: streams[e["s"]].append(e)
print("span %.1f us, %d events, streams: %s" % (span, len(ev), {k: len(v) for k, v in streams.items()}))
for s, L in sorted(streams.items(), key=lambda kv: -len(kv[1])):
    busy = sum(e["d"] for e in L)
    print("stream %s: %d kernels, busy %.1f us (%.0f%% of span), first %.1f last %.1f" % (s, len(L), busy, 100 * busy / span, L[0]["t"], L[-1]["t"] + L[-1]["d"]))
# chain stream = the one with most kernels
chain = max(streams.values(), key=len)
gaps = []
for a, b in zip(chain, chain[1:]):
    g = b["t"] - (a["t"] + a["d"])
    gaps.append((g, a, b))
tot_gap = sum(g for g, _, _ in gaps if g > 0)
print("chain: kernel time %.1f us, gaps %.1f us (%d gaps > 5us: %.1f us)" % (sum(e["d"] for e in chain), tot_gap, sum(1 for g, _, _ in gaps if g > 5), sum(g for g, _, _ in gaps if g > 5)))
hist = collections.Counter()
for g, _, _ in gaps: hist[min(int(max(g, 0)), 20)] += 1
print("gap histogram (us -> count):", sorted(hist.items()))
print("largest gaps:")
for g, a, b in sorted(gaps, key=lambda x: -x[0])[:25]:
    print("  %.1f us after %s (%.1f us) before %s @%.1f" % (g, a["n"][:40], a["d"], b["n"][:40], b["t"]))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in chain:
    k = (e["n"][:44], str(e["g"]))
    agg[k][0] += 1; agg[k][1] += e["d"]
print("chain kernels by (name, grid):")
for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("  %-44s %-16s n=%4d tot=%8.1f avg=%6.1f" % (k[0], k[1], n, d, d / n))
