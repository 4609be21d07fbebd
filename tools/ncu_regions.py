#!/usr/bin/env python
"""Group a kernel's SASS (ncu source page CSV) into runs of similar execution count: shows where warp instructions go.
usage: tools/ncu_regions.py src.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = rows[2:]
tot = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
print('total warp instructions %.2fG' % (tot / 1e9))
seg = []; cur = None
for n, r in enumerate(data):
    c = int(r[ix['Instructions Executed']] or 0); t = float(r[ix['Avg. Threads Executed']] or 0)
    key = round(c / (tot / 3000.0))
    if cur and abs(cur[2] - key) <= max(1, 0.08 * key): cur[1] = n; cur[3] += c; cur[4] += t * c
    else:
        cur = [n, n, key, c, t * c]; seg.append(cur)
for s in seg:
    if s[3] > tot * minpct / 100:
        print(f"rows {s[0]:5d}-{s[1]:5d} n={s[1]-s[0]+1:4d} exec/instr={s[3]/(s[1]-s[0]+1)/1e6:8.1f}M total={s[3]/1e9:6.2f}G ({100*s[3]/tot:4.1f}%) avgthr={s[4]/max(s[3],1):5.1f}  first: {data[s[0]][ix['Source']].strip()[:60]}")
