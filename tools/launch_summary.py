#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals of the last N launches."""
import collections, csv, sys
path, n = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(path)))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        hdr, start = r, i
        break
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
seq = [(r[ki].split('(')[0][-34:], float(r[vi]), r[ui]) for r in rows[start + 1:] if len(r) > vi]
seq = [(k, v / 1000 if u in ('ns', 'nsecond') else v) for k, v, u in seq]
agg = collections.OrderedDict()
for k, v in seq[-n:]:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, v) in agg.items():
    print("%-38s x%-3d %9.1f us" % (k, c, v))
print("sum %.1f us over %d launches" % (sum(v for c, v in agg.values()), n))
