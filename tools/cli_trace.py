#!/usr/bin/env python
"""Diagnostic: set-up / tear-down trace of one CLI run on a small lane (MGK_TRACE=1).  python tools/cli_trace.py [pairs]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
d = "/dev/shm/mgk_trace"
try:
    r1, r2, sheet, _ = file_bench.write_lane_blocks(synth.CONFIG2, n, d + "/in")
    a = lanes.demux_args(synth.CONFIG2, r1, r2, sheet, d + "/out")
    a[a.index("-t") + 1] = str(os.cpu_count())
    for k in range(2):
        t = time.perf_counter()
        p = subprocess.run([CLI_BIN] + a, capture_output=True, text=True, env=dict(os.environ, MGK_TRACE="1"))
        print("run", k, "wall %.2f s" % (time.perf_counter() - t), "rc", p.returncode)
        print(p.stderr[-1200:])
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"cli_trace_run{k}.txt"), "w") as f:
            f.write(p.stderr)
        print("\n".join(line[30:] for line in p.stdout.split("\n") if "after" in line or "Exection" in line or "pump waited" in line))
finally:
    shutil.rmtree(d, ignore_errors=True)
