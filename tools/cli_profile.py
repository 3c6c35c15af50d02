#!/usr/bin/env python
"""Diagnostic: the drop-in CLI on a multi-member .fq.gz lane under a few host-side settings, with the pump's own
wait accounting (where the wall time of a file-level run goes).  python tools/cli_profile.py [pairs]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = "/dev/shm/mgk_profile"
os.makedirs(d, exist_ok=True)
try:
    r1, r2, sheet, gs = file_bench.write_lane_blocks(synth.CONFIG2, n, d + "/in")
    print("lane written in %.1f s" % gs)
    variants = [[], ["--slots", "6"], ["--chunk-size", "33554432"], ["--reader-threads", "6"], ["--reader-threads", "12"], ["--host-gzip"]]
    for extra in variants:
        a = lanes.demux_args(synth.CONFIG2, r1, r2, sheet, d + "/out", extra)
        a[a.index("-t") + 1] = str(os.cpu_count())
        t = time.perf_counter()
        p = subprocess.run([CLI_BIN] + a, capture_output=True, text=True)
        dt = time.perf_counter() - t
        print(extra, "rc", p.returncode, round(n / dt / 1e6, 3), "M pairs/s", round(dt, 2), "s")
        for line in p.stdout.split("\n"):
            if "after" in line or "pump waited" in line or "Hot path" in line:
                print("    ", line[30:])
        if p.returncode:
            print(p.stderr[-500:])
finally:
    shutil.rmtree(d, ignore_errors=True)
