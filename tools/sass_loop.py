#!/usr/bin/env python
"""Histogram the opcodes of a kernel's SASS between two addresses (default: the innermost backward
branch's target .. the branch), to count FP64 / MUFU / other instructions per loop iteration.
usage: cuobjdump -sass -fun <mangled> file.o | python tools/sass_loop.py [lo hi]"""
import re, sys, collections
rows = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)[^;]*?(0x[0-9a-f]+)?\s*;", line)
    if m:
        rows.append((int(m.group(1), 16), m.group(2), line))
if len(sys.argv) >= 3:
    lo, hi = int(sys.argv[1], 16), int(sys.argv[2], 16)
else:
    best = None
    for a, op, line in rows:
        if op == "BRA":
            t = re.search(r"(0x[0-9a-f]+)\s*;", line)
            if t:
                tgt = int(t.group(1), 16)
                if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                    best = (tgt, a)
    lo, hi = best
c = collections.Counter(op for a, op, _ in rows if lo <= a <= hi)
tot = sum(c.values())
fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"loop 0x{lo:x}..0x{hi:x}: {tot} instructions, {fp64} FP64, {c.get('MUFU',0)} MUFU")
print(" ".join(f"{k}:{v}" for k, v in c.most_common()))
