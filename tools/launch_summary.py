"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ki, vi = H.index('Kernel Name'), H.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 2:]:
    if len(r) <= vi:
        continue
    n = r[ki].split('(')[0][-44:]
    agg[n][0] += 1
    agg[n][1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:46s} {v[0]:4d} {v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0] / 1e3:.1f}")
print("total %.3f ms" % (tot / 1e6))
