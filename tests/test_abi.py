"""The C-ABI library loads and exports every symbol include/mgikit_gpu.h declares.
No compute calls here (no GPU in this container); on a CPU-only box mgk_create must
refuse loudly instead of falling back."""
import ctypes
import os
import re

import pytest

from mgikit_b200 import synth
from mgikit_b200.capi import ABI_SYMBOLS, MgkError, load_gpu_library
from util import make_hotpath

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_and_binding_agree():
    with open(os.path.join(ROOT, "include", "mgikit_gpu.h")) as f:
        declared = set(re.findall(r"\b(mgk_[a-z_0-9]+)\s*\(", f.read()))
    assert declared == {"mgk_" + s for s in ABI_SYMBOLS}


def test_gpu_library_exports_every_declared_symbol():
    lib = load_gpu_library()
    for s in ABI_SYMBOLS:
        assert hasattr(lib, "mgk_" + s), f"libmgikit_gpu.so does not export mgk_{s}"


def test_oracle_mirrors_the_abi(oracle_lib):
    for s in ["create", "destroy", "last_error", "submit", "collect", "release", "counters", "reset_counters",
              "allreduce_counters", "reduced_counters", "host_alloc", "host_free", "scan_newlines", "match_barcodes"]:
        assert hasattr(oracle_lib, "orc_" + s)


def _has_cuda_device():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda_device(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_create_fails_loudly():
    lib = load_gpu_library()
    with pytest.raises(MgkError) as e:
        make_hotpath(lib, "mgk_", synth.run_spec(synth.CONFIG2))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_links_the_oracle():
    """The product binaries must not depend on anything under oracle/."""
    import subprocess
    for p in ("mgikit_b200/lib/libmgikit_gpu.so", "mgikit_b200/bin/mgikit"):
        full = os.path.join(ROOT, p)
        out = subprocess.run(["nm", "-D", "--undefined-only", full], capture_output=True, text=True).stdout
        out += subprocess.run(["ldd", full], capture_output=True, text=True).stdout
        assert "orc_" not in out and "emu_" not in out and "liboracle" not in out
