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
    print("\n".join(lines[:30]))
    print("...")
    print("\n".join(lines[-12:]))
    print("\n".join(line[30:] for line in p.stdout.split("\n") if "pump" in line or "ready after" in line))
finally:
    shutil.rmtree(d, ignore_errors=True)
