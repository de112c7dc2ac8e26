#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
Usage: launch_summary.py <launches.csv> <out.md> [title]"""
import csv
import sys
from collections import defaultdict

src, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else src
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    tot[r[ki]] += float(r[vi].replace(",", "")) / 1e6
    cnt[r[ki]] += 1
total = sum(tot.values())
with open(out, "w") as fh:
    fh.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (raw list: {src.split('/')[-1]}). "
             "Times are cold-cache and serialised: compare SHARES.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k in sorted(tot, key=tot.get, reverse=True):
        fh.write(f"| `{k[:70]}` | {cnt[k]} | {tot[k]:.3f} | {100 * tot[k] / total:.2f}% |\n")
print("wrote", out)
