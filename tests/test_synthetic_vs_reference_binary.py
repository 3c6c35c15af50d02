"""Seeded synthetic lanes shaped like BASELINE.json configs 2-5: outputs must equal
what the REFERENCE'S OWN BINARY produced (tests/golden/synthetic_reference.json, made
by make_golden.py; when oracle/_ref/mgikit is present the reference is re-run live)."""
import gzip
import hashlib
import json
import os
import shutil
import tempfile

import pytest

import lanes
from golden.make_golden import synthetic_variants

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "synthetic_reference.json")) as f:
    GOLD = json.load(f)
VARIANTS = synthetic_variants()
REF_BIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "mgikit")


def _check(binary, name, lane, extra, n, extra_cli=()):
    d = tempfile.mkdtemp(prefix="lane_")
    try:
        r1, r2, sheet = lanes.write_lane(lane, n, d)
        for p in (r1, r2):
            if p:  # the generator must still produce the bytes the goldens were made from
                with gzip.open(p, "rb") as f:
                    assert hashlib.sha256(f.read()).hexdigest() == GOLD[name]["inputs"][os.path.basename(p)], \
                        "synthetic generator drifted: regenerate tests/golden"
        out = os.path.join(d, "out")
        lanes.run_binary(binary, lanes.demux_args(lane, r1, r2, sheet, out, list(extra) + list(extra_cli)))
        got = lanes.digest_dir(out)
        want = GOLD[name]["files"]
        assert sorted(got) == sorted(want), f"{name}: file sets differ: {sorted(set(got) ^ set(want))[:6]}"
        diff = [k for k in want if got[k] != want[k]]
        assert not diff, f"{name}: {len(diff)} files differ from the reference binary's output, e.g. {diff[:4]}"
    finally:
        shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("name,lane,extra,n", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_oracle_cli_equals_reference_binary(oracle_cli, name, lane, extra, n):
    _check(oracle_cli, name, lane, extra, n)


def test_device_logic_on_cpu_multi_chunk_equals_reference_binary(emu_cli):
    """The per-record device logic through the host pump with many chunks in flight (chunk cuts, per-sample order)."""
    for k in (0, 1, 3):
        name, lane, extra, n = VARIANTS[k]
        _check(emu_cli, name, lane, extra, n, extra_cli=("--chunk-size", "524288"))


@pytest.mark.gpu
@pytest.mark.parametrize("name,lane,extra,n", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_cuda_cli_equals_reference_binary(product_cli, name, lane, extra, n):
    _check(product_cli, name, lane, extra, n)


@pytest.mark.gpu
def test_cuda_cli_multi_chunk_equals_reference_binary(product_cli):
    name, lane, extra, n = VARIANTS[0]
    _check(product_cli, name, lane, extra, n, extra_cli=("--chunk-size", "1048576"))


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference binary not available")
def test_goldens_are_what_the_reference_binary_prints_now():
    name, lane, extra, n = VARIANTS[-1]
    _check(REF_BIN, name, lane, extra, n)
