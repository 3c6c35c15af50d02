#!/usr/bin/env python
"""Diagnostic: device time of the gzip kernels (MGK_OUT_GZIP) on resident chunks of a lane, next to the text path.
python tools/gzip_timing.py [2|3|4|5] [pairs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from mgikit_b200 import synth
from mgikit_b200.capi import MGK_IN_DEVICE, MGK_OUT_DEVICE, MGK_OUT_GZIP, load_gpu_library
from util import device_copy, make_hotpath

which = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
lane = {2: synth.CONFIG2, 3: synth.CONFIG3, 4: synth.CONFIG4, 5: synth.CONFIG5}[which]
r1, r2 = synth.make_block(lane, 0, n)
spec = synth.run_spec(lane, max_chunk_bytes=max(r1.size if r1 is not None else 0, r2.size) + 4096, max_chunk_records=n + 64, n_slots=2)
hp = make_hotpath(load_gpu_library(), "mgk_", spec)
d2, a2 = device_copy(r2)
d1, a1 = device_copy(r1) if r1 is not None and r1.size else (None, None)
rows, gz = [], []
for k in range(6):
    hp.submit(k, a2, a1, MGK_IN_DEVICE | MGK_OUT_DEVICE | MGK_OUT_GZIP)
    res = hp.collect(device_out=True)
    rows.append(hp.last_timings())
    gz.append(hp.last_gzip_ms())
    hp.release(k)
t = np.array(rows[2:]).mean(axis=0)
g = float(np.mean(gz[2:]))
gz_bytes = res.out_r2_bytes + res.out_r1_bytes
hp.submit(9, a2, a1, MGK_IN_DEVICE | MGK_OUT_DEVICE)
text = hp.collect(device_out=True)
hp.release(9)
text_bytes = text.out_r2_bytes + text.out_r1_bytes
print(f"{lane.name}: {n} units, text out {text_bytes} B, gzip out {gz_bytes} B (ratio {gz_bytes / text_bytes:.3f})")
print(f"demultiplex kernels {t[4]:.3f} ms, gzip kernels {g:.3f} ms = {text_bytes / (g * 1e-3) / 1e9:.0f} GB/s of text, {n / (g * 1e-3) / 1e6:.0f} M units/s")
hp.close()
