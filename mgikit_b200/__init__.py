"""mgikit_b200 — B200-native `mgikit demultiplex` hot path.

The product is native code: ``lib/libmgikit_gpu.so`` (sm_100a kernels behind the
C ABI of ``include/mgikit_gpu.h``) and ``bin/mgikit`` (the drop-in CLI).  This
Python package is only the thin ctypes binding used by the tests and by
``bench.py``; it never computes anything itself and has no CPU fallback.
"""
from .capi import (HotPath, Result, load_gpu_library, MgkConfig, MgkTemplate, MgkResult,  # noqa: F401
                   MGK_IN_DEVICE, MGK_OUT_DEVICE, MgkError)
from .plan import Sample, RunSpec, build_config  # noqa: F401

__all__ = ["HotPath", "Result", "load_gpu_library", "Sample", "RunSpec", "build_config"]
