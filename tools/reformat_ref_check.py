#!/usr/bin/env python
"""Diagnostic: the REFERENCE binary's `reformat` on the block-written one-sample lane of the file-level bench, under a
timeout, log tail printed (it took 16 s for 2 M pairs in the build container and did not finish in 240 s on a GPU box)."""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
d = "/dev/shm/mgk_rfref"
try:
    f1, f2 = file_bench.write_reformat_lane(synth.CONFIG2, n, d + "/in")
    a = ["reformat", "-f", f1, "-r", f2, "--lane", "L01", "--run", "20240101", "--instrument", "INS", "--force", "--umi-length", "8",
         "--barcode", synth.REFORMAT_BARCODE, "-o", d + "/out"]
    t = time.perf_counter()
    try:
        p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "mgikit")] + a, capture_output=True, text=True, timeout=40)
        print("rc", p.returncode, "wall %.2f s for %d pairs" % (time.perf_counter() - t, n))
        print(p.stdout[-900:])
    except subprocess.TimeoutExpired as e:
        print("TIMEOUT after 40 s")
        print((e.stdout or b"")[-1500:].decode(errors="replace"))
        print((e.stderr or b"")[-600:].decode(errors="replace"))
    os.system("free -g | head -2; nproc; ls -la %s/out 2>/dev/null | head" % d)
finally:
    shutil.rmtree(d, ignore_errors=True)
