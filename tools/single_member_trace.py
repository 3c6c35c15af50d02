#!/usr/bin/env python
"""Diagnostic: MGK_TRACE of the CLI on a single-member lane (part timeline of both inputs)."""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import file_bench
import lanes
from mgikit_b200 import synth
from mgikit_b200.capi import CLI_BIN

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = "/dev/shm/mgk_single"
try:
    r1, r2, sheet, _ = file_bench.write_lane_blocks(synth.CONFIG2, n, d + "/in", single_member=True)
    a = lanes.demux_args(synth.CONFIG2, r1, r2, sheet, d + "/out")
    a[a.index("-t") + 1] = str(os.cpu_count())
    t = time.perf_counter()
    p = subprocess.run([CLI_BIN] + a, capture_output=True, text=True, env=dict(os.environ, MGK_TRACE="1"))
    print("wall %.2f" % (time.perf_counter() - t))
    lines = [l for l in p.stderr.split("\n") if "gz trace" in l or "create" in l]
    import re
    slow = [l for l in lines if "decoded" in l and float(re.search(r"in ([0-9.]+) ms", l).group(1)) > 100]
    print("parts slower than 100 ms:", len(slow))
    print("\n".join(slow[:10]))
    walk = [l for l in lines if "walker" in l]
    print("\n".join(walk[::max(1, len(walk) // 30)]))
    print("\n".join(lines[-6:]))
    print("\n".join(line[30:] for line in p.stdout.split("\n") if "pump" in line or "ready after" in line))
finally:
    shutil.rmtree(d, ignore_errors=True)
