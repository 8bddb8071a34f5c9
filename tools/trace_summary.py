"""Summarise a task trace of the factorization kernel (MIPM_TASK_TRACE, written by mipm_ls_factorize_profile)."""
import csv
import statistics
import sys

names = {0: 'EA', 1: 'DIAG', 2: 'PANEL', 3: 'TRAIL', 4: 'LEAF'}
rows = list(csv.DictReader(open(sys.argv[1])))
for r in rows:
    for k in r:
        r[k] = float(r[k]) if '.' in r[k] else int(r[k])
fronts = [int(a) for a in sys.argv[2:]] or [rows[-1]['front']]
for fr in fronts:
    print("front", fr)
    for r in rows:
        if r['front'] == fr and r['type'] != 0:
            print(f"  t{r['task']} {names[r['type']]:5s} a={r['a']} b={r['b']} c={r['c']} d={r['d']} need={r['need']} "
                  f"wait={r['wait_us']:.1f} start={r['start_us']:.1f} end={r['end_us']:.1f} dur={r['end_us'] - r['start_us']:.1f} sm={r['sm']}")
    ea = [r for r in rows if r['front'] == fr and r['type'] == 0]
    if ea:
        print(f"  EA x{len(ea)}: start {min(r['start_us'] for r in ea):.1f} end {max(r['end_us'] for r in ea):.1f}")
for t in range(5):
    d = [r['end_us'] - r['start_us'] for r in rows if r['type'] == t]
    if d:
        print(names[t], len(d), "mean %.1f median %.1f max %.1f sum %.0f" % (statistics.mean(d), statistics.median(d), max(d), sum(d)))
# durations by K length for PANEL/TRAIL/DIAG
import collections
by = collections.defaultdict(list)
for r in rows:
    if r['type'] == 1: by[('DIAG', r['a'] - r['b'])].append(r['end_us'] - r['start_us'])
    if r['type'] == 2: by[('PANEL', r['a'] - r['b'])].append(r['end_us'] - r['start_us'])
    if r['type'] == 3: by[('TRAIL', 16 * ((r['b'] - r['a'] + 15) // 16))].append(r['end_us'] - r['start_us'])
for k in sorted(by):
    print(k, len(by[k]), "median %.1f min %.1f" % (statistics.median(by[k]), min(by[k])))
# concurrency over time
ev = []
for r in rows:
    ev.append((r['start_us'], 1)); ev.append((r['end_us'], -1))
ev.sort()
T = max(r['end_us'] for r in rows)
nb = 30
busy = [0.0] * nb
cur, last = 0, 0.0
for t, d in ev:
    b0 = int(last / T * nb); b1 = int(min(t, T - 1e-9) / T * nb)
    for b in range(b0, b1 + 1):
        lo = max(last, b * T / nb); hi = min(t, (b + 1) * T / nb)
        if hi > lo: busy[b] += cur * (hi - lo)
    cur += d; last = t
print("avg running tasks per %.0f us slice:" % (T / nb), " ".join("%d" % (b / (T / nb)) for b in busy))
