#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel / grid."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix['Metric Name']] != 'gpu__time_duration.sum': continue
    k = r[ix['Kernel Name']]; v = float(r[ix['Metric Value']].replace(',', '')) * 1e-3
    agg[(k[:72], r[ix['Grid Size']], r[ix['Block Size']])][0] += 1
    agg[(k[:72], r[ix['Grid Size']], r[ix['Block Size']])][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
for g, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:4d} x {t/n:8.2f} us = {t:9.1f} us {100*t/tot:5.1f}%  grid={g[1]} block={g[2]}  {g[0]}")
