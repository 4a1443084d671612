"""Per (kernel, grid, block) totals of an ncu launch list: python scripts/launch_groups.py gpurun_out/launches.csv [top]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    if r['Metric Name'] != 'gpu__time_duration.sum': continue
    k = r['Kernel Name'].split('(')[0].replace('<unnamed>::', '')[-44:]
    v = float(r['Metric Value'])
    if r['Metric Unit'] in ('us', 'usecond'): v *= 1e3
    a = agg.setdefault((k, r['Grid Size'], r['Block Size']), [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("total %.1f us over %d launches" % (tot / 1e3, sum(a[0] for a in agg.values())))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-44s %-15s %-13s n=%3d tot=%7.1fus avg=%6.1fus" % (key[0], key[1], key[2], a[0], a[1] / 1e3, a[1] / a[0] / 1e3))
