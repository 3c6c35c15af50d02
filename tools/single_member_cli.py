#!/usr/bin/env python
"""Diagnostic: the CLI and the reference on a lane written as ONE gzip member per file.  python tools/single_member_cli.py [pairs]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = "/dev/shm/mgk_single"
try:
    t = time.time()
    r1, r2, sheet, _ = file_bench.write_lane_blocks(synth.CONFIG2, n, d + "/in", single_member=True)
    print("lane written in %.1f s, %d bytes" % (time.time() - t, os.path.getsize(r2)))
    cores = os.cpu_count()
    for extra in ([], ["--reader-threads", str(cores)], ["--reader-threads", "1"]):
        a = lanes.demux_args(synth.CONFIG2, r1, r2, sheet, d + "/out", extra)
        a[a.index("-t") + 1] = str(cores)
        t = time.perf_counter()
        p = subprocess.run([CLI_BIN] + a, capture_output=True, text=True)
        dt = time.perf_counter() - t
        print(extra, p.returncode, round(dt, 2), "s", round(n / dt / 1e6, 3), "M pairs/s")
        print("\n".join(line[30:] for line in p.stdout.split("\n") if "pump" in line or "ready after" in line))
        print(p.stderr[-300:])
    a = lanes.demux_args(synth.CONFIG2, r1, r2, sheet, d + "/outref")
    a[a.index("-t") + 1] = str(cores)
    t = time.perf_counter()
    p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "mgikit")] + a, capture_output=True, text=True)
    dt = time.perf_counter() - t
    print("reference", p.returncode, round(dt, 2), "s", round(n / dt / 1e6, 3), "M pairs/s")
    print("reports equal", file_bench._digest_dir(d + "/out", False) == file_bench._digest_dir(d + "/outref", False))
finally:
    shutil.rmtree(d, ignore_errors=True)
