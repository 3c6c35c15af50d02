#!/usr/bin/env python
"""Diagnostic: the product CLI's `reformat` on a block-written one-sample lane, under a timeout, log tail printed.
python tools/reformat_file_check.py [pairs] [extra CLI flags...]"""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import file_bench
from mgikit_b200 import synth
from mgikit_b200.capi import CLI_BIN

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
extra = sys.argv[2:]
d = "/dev/shm/mgk_rfcheck"
try:
    f1, f2 = file_bench.write_reformat_lane(synth.CONFIG2, n, d + "/in")
    a = ["reformat", "-f", f1, "-r", f2, "--lane", "L01", "--run", "20240101", "--instrument", "INS", "--force", "--umi-length", "8",
         "--barcode", synth.REFORMAT_BARCODE, "-o", d + "/out", "-t", str(os.cpu_count())] + extra
    t = time.perf_counter()
    try:
        p = subprocess.run([CLI_BIN] + a, capture_output=True, text=True, timeout=60)
        print("rc", p.returncode, "wall %.2f s" % (time.perf_counter() - t))
        print(p.stdout[-1500:])
        print(p.stderr[-800:])
    except subprocess.TimeoutExpired as e:
        print("TIMEOUT after 60 s")
        print((e.stdout or b"")[-2500:].decode(errors="replace"))
        print((e.stderr or b"")[-1200:].decode(errors="replace"))
finally:
    shutil.rmtree(d, ignore_errors=True)
