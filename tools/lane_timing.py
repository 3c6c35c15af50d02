#!/usr/bin/env python
"""Diagnostic (not the bench contract): per-kernel CUDA-event times of one resident chunk of a synthetic lane,
for the BASELINE configs that are parity cases rather than bench lines.  python tools/lane_timing.py [2|3|4|5] [pairs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from mgikit_b200 import synth
from mgikit_b200.capi import MGK_IN_DEVICE, MGK_OUT_DEVICE, load_gpu_library
from util import device_copy, make_hotpath

which = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
lane = {2: synth.CONFIG2, 3: synth.CONFIG3, 4: synth.CONFIG4, 5: synth.CONFIG5}[which]
r1, r2 = synth.make_block(lane, 0, n)
spec = synth.run_spec(lane, max_chunk_bytes=max(r1.size if r1 is not None else 0, r2.size) + 4096, max_chunk_records=n + 64, n_slots=2)
hp = make_hotpath(load_gpu_library(), "mgk_", spec)
d2, a2 = device_copy(r2)
d1, a1 = device_copy(r1) if r1 is not None and r1.size else (None, None)
rows = []
for k in range(8):
    hp.submit(k, a2, a1, MGK_IN_DEVICE | MGK_OUT_DEVICE)
    res = hp.collect(device_out=True)
    rows.append(hp.last_timings())
    hp.release(k)
t = np.array(rows[3:]).mean(axis=0)
out_bytes = res.out_r2_bytes + res.out_r1_bytes
in_bytes = r2.size + (r1.size if r1 is not None else 0)
names = ["scan_newlines", "match_plan", "bin_scan", "emit_qc", "chunk"]
print(lane.name, n, "pairs;", in_bytes // n, "B in +", out_bytes // n, "B out per unit")
print({k: round(float(v), 4) for k, v in zip(names, t)})
print("emit: %.0f GB/s algorithmic; chunk: %.1f M units/s" % ((in_bytes + out_bytes) / (t[3] * 1e-3) / 1e9, n / (t[4] * 1e-3) / 1e6))
hp.close()
