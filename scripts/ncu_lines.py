#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python scripts/ncu_lines.py src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = collections.OrderedDict()


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


kern = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].isdigit():
        key = (cur_file, int(r[0]), r[1].strip()[:100])
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += num(r[hdr.index("Instructions Executed")])
        a[1] += num(r[hdr.index("# Samples")])
        a[2] += num(r[hdr.index("Thread Instructions Executed")])
tot = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print("total warp-instructions", tot, "samples", ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / ts * 100:5.1f}% samp {v[2] / max(v[0], 1):5.1f} lanes  {k[0]}:{k[1]}  {k[2]}")
