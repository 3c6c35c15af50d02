#!/usr/bin/env python
"""Per-function totals (warp-instructions, stall samples) from `ncu --page source --print-source cuda,sass --csv`,
attributing every source line of a .cuh file to the last MGK_HD function that starts above it."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
root = sys.argv[3] if len(sys.argv) > 3 else "mgikit_b200/csrc/gpu"
cache = {}


def fn(file, line):
    if file not in cache:
        try:
            src = open(f"{root}/{file}").read().split("\n")
        except OSError:
            src = []
        fs = []
        for i, l in enumerate(src, 1):
            m = re.match(r"^(?:MGK_HD|__global__|__device__|inline|template).*?\b(\w+)\(", l)
            if m and not l.startswith("template"):
                fs.append((i, m.group(1)))
        cache[file] = fs
    name = "?"
    for s, n in cache[file]:
        if s <= line:
            name = n
    return name


hdr = None
cur = ""
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and r and r[0].isdigit():
        try:
            ie = int(r[hdr.index("Instructions Executed")])
            sm = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        a = agg.setdefault((cur, fn(cur, int(r[0]))), [0, 0])
        a[0] += ie
        a[1] += sm
tot = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0] / units:10.1f} instr/unit {100 * v[0] / tot:5.1f}%   samples {100 * v[1] / max(ts, 1):5.1f}%  {k[0]}:{k[1]}")
