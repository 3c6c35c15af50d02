#!/usr/bin/env python
"""Split `ncu --page source --print-source sass --csv` of one kernel into regions delimited by BAR.SYNC / RET / CALL
targets and print warp-instructions, stall samples and the opcode mix per region."""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iA, iS, iN, iI = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
regions = []
cur = {"start": 0, "instr": 0, "samples": 0, "ops": collections.Counter(), "n": 0, "first": ""}
for k, r in enumerate(rows[2:]):
    if len(r) <= iI:
        continue
    op = r[iS].split()
    op = [t for t in op if not t.startswith("@")]
    name = op[0].split(".")[0] if op else "?"
    ie, sm = int(r[iI]), int(r[iN])
    cur["instr"] += ie
    cur["samples"] += sm
    cur["ops"][name] += ie
    cur["n"] += 1
    if not cur["first"]:
        cur["first"] = r[iS].strip()
    if name in ("BAR", "RET", "EXIT") or (name == "BRA" and ie == 0 and False):
        cur["end"] = k
        regions.append(cur)
        cur = {"start": k + 1, "instr": 0, "samples": 0, "ops": collections.Counter(), "n": 0, "first": ""}
cur["end"] = len(rows)
regions.append(cur)
tot = sum(r["instr"] for r in regions)
ts = sum(r["samples"] for r in regions)
print("total warp-instructions", tot, "samples", ts)
for r in regions:
    if r["instr"] == 0:
        continue
    top = ", ".join(f"{o} {100 * v / r['instr']:.0f}%" for o, v in r["ops"].most_common(8))
    print(f"sass[{r['start']:5d}..{r['end']:5d}] static {r['n']:5d}  instr {r['instr']:11d} {100 * r['instr'] / tot:5.1f}%  samples {100 * r['samples'] / ts:5.1f}%  | {top}")
