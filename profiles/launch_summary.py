#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (tools/profile_round.sh: --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum --csv) next to the CUDA-event times of the same command without ncu.
usage: launch_summary.py launches.csv bench_plain.json [pairs_per_launch]"""
import csv
import json
import sys
from collections import OrderedDict

path, bench = sys.argv[1], sys.argv[2]
pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 2097152
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
head, rows = rows[0], rows[1:]
ix = {k: head.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launch = OrderedDict()
for r in rows:
    d = launch.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]].split("(")[0]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[r[ix["Metric Name"]]] = v * scale
k = OrderedDict()
for d in launch.values():
    e = k.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    e["n"] += 1
    e["us"] += d.get("gpu__time_duration.sum", 0.0)
    e["rd"] += d.get("dram__bytes_read.sum", 0.0)
    e["wr"] += d.get("dram__bytes_write.sum", 0.0)
chunk = sum(e["us"] / e["n"] for e in k.values())
print(f"{'kernel':<24}{'launches':>9}{'mean us':>10}{'share':>8}{'DRAM read B/pair':>18}{'DRAM written B/pair':>21}")
tr = tw = 0.0
for name, e in k.items():
    us = e["us"] / e["n"]
    rd, wr = e["rd"] / e["n"] / pairs, e["wr"] / e["n"] / pairs
    tr += rd
    tw += wr
    print(f"{name:<24}{e['n']:>9}{us:>10.1f}{100 * us / chunk:>7.1f}%{rd:>18.1f}{wr:>21.1f}")
print(f"{'whole chunk':<24}{'':>9}{chunk:>10.1f}{'':>8}{tr:>18.1f}{tw:>21.1f}")
b = json.loads(open(bench).read().strip().splitlines()[-1])
alg = b["roofline"]["algorithmic_bytes_per_launch"] / pairs
print(f"\nDRAM traffic of the whole path: {tr + tw:.0f} B per pair against {alg:.0f} B algorithmic (every input byte read once, "
      f"every output byte written once) = {(tr + tw) / alg:.2f}x")
ms = b["roofline"]["kernel_ms"]
tot = sum(v for kk, v in ms.items() if kk != "chunk")
print("\nbench.py kernel_ms of the same build, same command without ncu (CUDA events, one chunk in flight):")
for kk, v in ms.items():
    if kk != "chunk":
        print(f"  {kk:<16}{v:>9.4f} ms{100 * v / tot:>7.1f}%")
print(f"  value {b['value'] / 1e9:.3f} G pairs/s, k_emit roofline frac {b['roofline']['frac']:.3f}")
e = k.get("k_emit")
if e:
    print(f"\nk_emit DRAM bytes per launch: {(e['rd'] + e['wr']) / e['n']:.0f}")
