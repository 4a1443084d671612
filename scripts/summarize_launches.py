"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).

usage: python scripts/summarize_launches.py gpurun_out/launches.csv [skip_first_n_launches] > profiles/rN_launches.md
The per-launch times under ncu are cold-cache and serialised: what must agree with bench.py is each kernel's SHARE.
"""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_]+)(<[^(]*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def main() -> None:
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"])
            if r.get("Metric Unit") in ("us", "usecond"):
                v *= 1e3
            rows.append((int(r["ID"]), short(r["Kernel Name"]), v, r["Grid Size"], r["Block Size"]))
    rows = [r for r in rows if r[0] >= skip]
    agg = collections.OrderedDict()
    for _, k, ns, grid, blk in rows:
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += ns; a[2] = min(a[2], ns); a[3] = max(a[3], ns)
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share | avg us | min us | max us |")
    print("|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% | %.1f | %.1f | %.1f |" % (k, a[0], a[1] / 1e6, 100 * a[1] / total, a[1] / a[0] / 1e3,
                                                                   a[2] / 1e3, a[3] / 1e3))
    print("\n%d launches, %.3f ms summed kernel time (serialised, cold cache, under ncu)" % (len(rows), total / 1e6))


if __name__ == "__main__":
    main()
