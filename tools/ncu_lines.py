#!/usr/bin/env python
"""Attribute a kernel's executed warp instructions (ncu source page CSV) to source lines through nvdisasm -g line info.
usage: tools/ncu_lines.py src.csv kernel.dis(nvdisasm -g -c of the cubin) kernel-substring [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = rows[2:]
pat = sys.argv[3]; top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(sys.argv[2]).read().split("\n")
on = False; cur = ("?", 0); stack = ""; per = []
for l in lines:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m: on = pat in m.group(1); continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); stack = m.group(3); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l): per.append((cur, stack, l.strip()))
print("sass instrs in dis:", len(per), "rows in ncu:", len(data))
n = min(len(per), len(data))
tot = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
agg = collections.defaultdict(lambda: [0, 0.0, 0, 0])
for i in range(n):
    c = int(data[i][ix['Instructions Executed']] or 0); t = float(data[i][ix['Avg. Threads Executed']] or 0)
    s = int(data[i][ix['# Samples']] or 0)
    a = agg[per[i][0]]; a[0] += c; a[1] += c * t; a[2] += 1; a[3] += s
stot = sum(a[3] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:22s}:{k[1]:4d}  n={a[2]:4d} warp-instr={a[0]/1e9:6.2f}G ({100*a[0]/tot:4.1f}%) avgthr={a[1]/max(a[0],1):5.1f} samples={100*a[3]/max(stot,1):4.1f}%")
