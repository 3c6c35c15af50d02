"""CPU gate for the CUDA path: the device logic (device_logic.cuh compiled by g++,
lanes looped) must agree with the oracle bit for bit on every chunk-level case.
The GPU parity tests (test_gpu_parity.py) run the same cases through the kernels."""
import pytest

import cases
from util import run_both

ALL = cases.all_cases()


@pytest.mark.parametrize("name,spec,chunks", ALL, ids=[c[0] for c in ALL])
def test_emulated_device_logic_matches_oracle(oracle_lib, emu_lib, name, spec, chunks):
    run_both(oracle_lib, "orc_", emu_lib, "emu_", spec, chunks, what=name)
