#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump (SASS rows) into regions: splits the instruction stream where the
executed count changes (region boundaries = branch targets with different trip counts) and prints, per region,
instructions, executed warp-instructions, mean active threads and stall samples.
usage: python tools/ncu_source_hot.py prof_source.csv[.gz] [min_share]"""
import csv
import gzip
import sys

path = sys.argv[1]
fh = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(fh))
hdr = rows[1]
ix = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
st_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    data.append((r[ix["Source"]].strip(), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]),
                 int(r[ix["Thread Instructions Executed"]]), [int(r[i]) for i in st_cols]))
tot_inst = sum(d[2] for d in data)
tot_samp = sum(d[1] for d in data)
print(f"{len(data)} SASS instructions, {tot_inst:.3e} warp-instructions executed, {tot_samp} stall samples")
regions = []
cur = None
for k, d in enumerate(data):
    ex = d[2]
    if cur is None or not (0.8 * cur["ex"] <= ex <= 1.25 * cur["ex"]):
        cur = {"start": k, "ex": max(ex, 1), "n": 0, "inst": 0, "thr": 0, "samp": 0, "st": [0] * len(st_cols), "ops": {}}
        regions.append(cur)
    cur["n"] += 1
    cur["inst"] += d[2]
    cur["thr"] += d[3]
    cur["samp"] += d[1]
    for i, v in enumerate(d[4]):
        cur["st"][i] += v
    parts = d[0].split()
    op = parts[0] if parts else "?"
    if op.startswith("@") and len(parts) > 1:
        op = parts[1]
    op = op.split(".")[0]
    cur["ops"][op] = cur["ops"].get(op, 0) + 1
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
names = [hdr[i] for i in st_cols]
for r in regions:
    if r["inst"] < minshare * tot_inst and r["samp"] < minshare * tot_samp:
        continue
    top = sorted(zip(r["st"], names), reverse=True)[:3]
    ops = sorted(r["ops"].items(), key=lambda kv: -kv[1])[:6]
    print(f"@{r['start']:5d} n={r['n']:5d} exec/inst={r['inst'] / r['n']:.3e} inst%={100 * r['inst'] / tot_inst:5.1f} "
          f"thr/inst={r['thr'] / max(r['inst'], 1):5.1f} samp%={100 * r['samp'] / max(tot_samp, 1):5.1f} "
          f"stalls={[(n, v) for v, n in top]} ops={ops}")
