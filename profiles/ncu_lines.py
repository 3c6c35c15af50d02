#!/usr/bin/env python
"""Aggregate `ncu --page source --print-source cuda,sass --csv` per CUDA source line:
warp-instructions executed and stall samples, top N lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
out = []
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and r and r[0].isdigit():
        ie = hdr.index("Instructions Executed")
        ss = hdr.index("# Samples")
        try:
            out.append((int(r[ie]), int(r[ss]), cur_file, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(o[0] for o in out)
tots = sum(o[1] for o in out)
print(f"total warp-instructions {tot}, samples {tots}")
byfile = {}
for o in out:
    byfile[o[2]] = byfile.get(o[2], 0) + o[0]
print(byfile)
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0]:>11d} {100.0 * o[0] / tot:5.1f}%  smp {100.0 * o[1] / max(tots, 1):4.1f}%  {o[2]}:{o[3]:<4d} {o[4]}")
