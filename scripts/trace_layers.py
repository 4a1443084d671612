"""Per-layer kernel times from the library's event profiler: run `SVAE_TRACE=1 SVAE_MULTI=0 python bench.py ... 2> trace.err`
(the profiled steps of bench.py print one TRACE line per launch with its geometry) and aggregate by (class, geometry).
usage: python scripts/trace_layers.py trace.err [steps_profiled]"""
import collections, re, sys
agg = collections.OrderedDict()
pat = re.compile(r"TRACE (\S+) B=(\d+) Hin=(\d+) Cin=(\d+) Hout=(\d+) Cout=(\d+) k=(\d+) s=(\d+) mode=(\d+) ms=([\d.]+)")
for line in open(sys.argv[1]):
    m = pat.search(line)
    if not m:
        continue
    key = (m.group(1),) + tuple(int(v) for v in m.groups()[1:9])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(m.group(10))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tot = sum(a[1] for a in agg.values())
print("total %.3f ms over %d launches (%.3f ms per step)" % (tot, sum(a[0] for a in agg.values()), tot / steps))
print("%-22s %4s %4s %5s %4s %5s %2s %2s %4s %6s %9s %8s %8s" % ("class", "B", "Hin", "Cin", "Hout", "Cout", "k", "s", "mode", "n/step", "ms/step", "us/launch", "TFLOP/s"))
for key, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    kc, B, Hin, Cin, Hout, Cout, k, s, mode = key
    pix = B * min(Hin * Hin, Hout * Hout) if k == 4 else B
    fl = 2.0 * pix * (16 if k == 4 else 1) * Cin * Cout
    print("%-22s %4d %4d %5d %4d %5d %2d %2d %4d %6.1f %9.3f %8.1f %8.1f" % (kc, B, Hin, Cin, Hout, Cout, k, s, mode, n / steps, ms / steps, 1e3 * ms / n,
                                                                       fl / (ms / n * 1e-3) / 1e12 if "gemm" in kc or "wgrad" in kc else 0.0))
