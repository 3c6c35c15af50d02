"""The reference's black-box `demultiplex` test matrix (tests/ref_matrix.py) replayed
against the committed golden bundle (tests/golden/, made by make_golden.py from the
reference's fixtures):
  * CPU (-m "not gpu"): the oracle CLI and the device-logic emulation CLI — this is
    what PINS the oracle to the reference's golden vectors;
  * GPU (-m gpu): the product CLI, i.e. the CUDA path end to end through files.
"""
import json
import os
import shutil
import tempfile

import pytest

import ref_matrix as M

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
ROOT_DATA = os.path.join(GOLDEN, "testing_data")
with open(os.path.join(GOLDEN, "reference_matrix.json")) as f:
    MATRIX = json.load(f)
CASES = M.all_cases()


def _run(binary, case, extra=()):
    out = tempfile.mkdtemp(prefix="mgk_out_")
    shutil.rmtree(out)
    try:
        p = M.run_case(binary, case, ROOT_DATA, out, extra)
        assert p.returncode == 0, f"{case.name}: exit {p.returncode}\n{p.stderr[-600:]}\n{p.stdout[-300:]}"
        M.check_stdout_shape(p.stdout)
        entry = MATRIX[case.name]
        listing = {k: tuple(v) for k, v in entry["files"].items()}
        M.compare_outputs(case, listing, out, entry.get("stats_override", ""))
    finally:
        shutil.rmtree(out, ignore_errors=True)


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_oracle_cli_reproduces_reference_fixture(oracle_cli, case):
    _run(oracle_cli, case)


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_device_logic_emulation_reproduces_reference_fixture(emu_cli, case):
    _run(emu_cli, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_cuda_cli_reproduces_reference_fixture(product_cli, case):
    _run(product_cli, case)


@pytest.mark.gpu
def test_cuda_cli_small_chunks_preserve_order(product_cli):
    """large_ds cut into many chunks: per-sample order must survive chunking and double buffering."""
    case = [c for c in CASES if c.name == "large-pe"][0]
    _run(product_cli, case, extra=("--chunk-size", "262144"))


def test_oracle_cli_small_chunks_preserve_order(oracle_cli):
    case = [c for c in CASES if c.name == "large-pe"][0]
    _run(oracle_cli, case, extra=("--chunk-size", "262144"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/testing_data"), reason="reference tree not mounted")
def test_golden_bundle_matches_the_reference_tree():
    """The committed digests are exactly what the reference's testing_data/expected holds."""
    for case in CASES:
        live = M.expected_listing(case, "/root/reference/testing_data")
        assert {k: list(v) for k, v in live.items()} == MATRIX[case.name]["files"], case.name
