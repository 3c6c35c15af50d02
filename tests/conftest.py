import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ORACLE_DIR = os.path.join(ROOT, "oracle", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _ensure_built(paths, target):
    if all(os.path.exists(p) for p in paths):
        return
    subprocess.run(["make", "-C", ROOT, target], check=True, capture_output=True)


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure)."""
    p = os.path.join(ORACLE_DIR, "liboracle.so")
    _ensure_built([p], "oracle")
    return ctypes.CDLL(p)


@pytest.fixture(scope="session")
def emu_lib():
    """CPU emulation of the device logic (test infrastructure)."""
    p = os.path.join(ORACLE_DIR, "libemu.so")
    _ensure_built([p], "oracle")
    return ctypes.CDLL(p)


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; a missing .so is an error, never a skip."""
    from mgikit_b200.capi import load_gpu_library
    return load_gpu_library()


@pytest.fixture(scope="session")
def oracle_cli():
    p = os.path.join(ORACLE_DIR, "mgikit_oracle")
    _ensure_built([p], "oracle")
    return p


@pytest.fixture(scope="session")
def emu_cli():
    p = os.path.join(ORACLE_DIR, "mgikit_emu")
    _ensure_built([p], "oracle")
    return p


@pytest.fixture(scope="session")
def product_cli():
    from mgikit_b200.capi import CLI_BIN
    assert os.path.exists(CLI_BIN), f"{CLI_BIN} is missing: run make"
    return CLI_BIN
