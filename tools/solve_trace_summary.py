"""Per-level summary of a MIPM_SOLVE_TRACE file (tools/profile_factor.py with PROFILE_OUT): when each elimination-tree
level finished in the forward and in the backward sweep, and the per-task times of one level on request.
Usage: python tools/solve_trace_summary.py trace.csv [level]"""
import collections
import csv
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
for r in rows:
    for k in r:
        r[k] = float(r[k])
lv = collections.defaultdict(list)
for r in rows:
    lv[(int(r["kind"]) in (0, 1, 4), int(r["level"]))].append(r)
prev = 0.0
for key in sorted(lv, key=lambda k: (not k[0], k[1] if k[0] else -k[1])):
    rs = lv[key]
    end = max(r["end_us"] for r in rs)
    dur = [r["end_us"] - r["start_us"] for r in rs]
    print("%s level %2d tasks %5d  first start %7.1f  last end %7.1f  (+%5.1f)  task time avg %5.1f max %5.1f"
          % ("fwd" if key[0] else "bwd", key[1], len(rs), min(r["start_us"] for r in rs), end, end - prev, sum(dur) / len(dur), max(dur)))
    prev = end
if len(sys.argv) > 2:
    L = int(sys.argv[2])
    for r in rows:
        if int(r["level"]) == L:
            print({k: (int(v) if k in ("task", "kind", "id", "level", "k", "r") else v) for k, v in r.items()})
