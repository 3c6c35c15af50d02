#!/usr/bin/env python
"""Diagnostic: one long gzip member through the reader (serial vs in parts), with the part trace.  python tools/gz_trace.py [pairs] [threads]"""
import ctypes
import gzip
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgikit_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 8
lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhostcodec.so"))
lib.hc_gunzip_file2.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]
path = "/dev/shm/gz_trace.fq.gz"
with gzip.open(path, "wb", compresslevel=1) as f:
    for b in range(n // 100000):
        f.write(synth.make_block(synth.CONFIG2, b, 100000)[1].tobytes())
size = n * 376
out = ctypes.create_string_buffer(size + 16)
for pmin, label in ((1 << 40, "serial"), (1 << 20, "in parts")):
    p, nl, m = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    err = ctypes.create_string_buffer(512)
    t = time.perf_counter()
    rc = lib.hc_gunzip_file2(path.encode(), threads, 1 << 30, pmin, 4 << 20, out, size, ctypes.byref(p), ctypes.byref(nl), ctypes.byref(m), err, 512)
    dt = time.perf_counter() - t
    print(label, "rc", rc, err.value.decode(), "%.2f s, %.0f MB/s" % (dt, p.value / dt / 1e6), file=sys.stderr)
os.remove(path)
