#!/usr/bin/env python
"""Decode the scheduling control bits (stall, write/read barrier, wait mask) of a kernel's SASS.
usage: tools/sass_ctrl.py <lib.so> <kernel-substring> [grep-regex]"""
import re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
flt = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
lines = txt.split("\n")
on = False
ins = []
i = 0
while i < len(lines):
    if "Function :" in lines[i]:
        on = pat in lines[i]
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i]) if on else None
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        if m2:
            w = (int(m2.group(1), 16) << 64) | int(m.group(3), 16)
            ins.append((m.group(1), m.group(2).strip(), (w >> 105) & 0xf, (w >> 109) & 1, (w >> 110) & 7, (w >> 113) & 7, (w >> 116) & 0x3f))
            i += 2
            continue
    i += 1
for a, s, stall, yld, wbar, rbar, wait in ins:
    if flt is None or flt.search(s):
        print(f"{a} {s[:70]:70s} stall={stall:2d} wbar={wbar if wbar<7 else '-'} rbar={rbar if rbar<7 else '-'} wait={wait:06b}")
