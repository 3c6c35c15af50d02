"""`mgikit reformat` (reference: src/lib.rs:2266-2464, the demultiplex == false branches of
process_buffer :1090-1147, :1240-1293, src/formater.rs, src/sample_data.rs:627-637).

The same hot path with one sample and the barcode taken from the header.  Held to (1) the reference's own 18 test
cases (tests/integration_test.rs:437-795) with its extras_test fixtures, (2) what the reference's binary wrote for 14
seeded one-sample lanes (tests/golden/reformat_reference.json; re-run live when the binary is present).
The reference's fourth command, `report`, has no parity target: its v2.2.0 binary panics on its own sample_stats
files (index out of bounds in write_reports), see test_reference_report_command_panics."""
import hashlib
import json
import os
import shutil
import subprocess
import tempfile

import pytest

import lanes
import reformat_cases as RF

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "golden", "testing_data")
with open(os.path.join(HERE, "golden", "reformat_matrix.json")) as f:
    MATRIX = json.load(f)
with open(os.path.join(HERE, "golden", "reformat_reference.json")) as f:
    GOLD = json.load(f)
FIXTURES = RF.fixture_cases()
LANES = RF.synthetic_lanes()
GPU_LANES = [l for l in LANES if not l.cpu_only]
REF_BIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "mgikit")


def _compare(name, got, want):
    assert sorted(got) == sorted(want), f"{name}: file sets differ: {sorted(set(got) ^ set(want))}"
    diff = [k for k in want if got[k] != want[k]]
    assert not diff, f"{name}: files differ from the reference's: {diff}"


def _fixture(binary, case, extra=()):
    d = tempfile.mkdtemp(prefix="rf_")
    try:
        out = os.path.join(d, "out")
        lanes.run_binary(binary, ["reformat"] + RF.resolve(case.args, DATA) + ["-o", out] + list(extra))
        want = MATRIX[case.name]["files"]
        if case.rename:
            want = {k.replace(*case.rename): v for k, v in want.items()}
        _compare(case.name, lanes.digest_dir(out), want)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _lane(binary, l, extra=()):
    d = tempfile.mkdtemp(prefix="rf_")
    try:
        t1, t2 = RF.lane_text(l)
        assert [hashlib.sha256(t).hexdigest() for t in (t1, t2) if t is not None] == GOLD[l.name]["inputs"], \
            "lane generator drifted: regenerate tests/golden (make_golden.py reformat)"
        p1, p2 = RF.write_rf_lane(l, d)
        out = os.path.join(d, "out")
        lanes.run_binary(binary, RF.rf_args(l, p1, p2, out) + list(extra))
        _compare(l.name, lanes.digest_dir(out), GOLD[l.name]["files"])
    finally:
        shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("case", FIXTURES, ids=[c.name for c in FIXTURES])
def test_oracle_cli_reproduces_reference_fixture(oracle_cli, case):
    _fixture(oracle_cli, case)


@pytest.mark.parametrize("l", LANES, ids=[l.name for l in LANES])
def test_oracle_cli_equals_reference_binary(oracle_cli, l):
    _lane(oracle_cli, l)


def test_oracle_cli_many_chunks(oracle_cli):
    _lane(oracle_cli, LANES[2], extra=("--chunk-size", "131072"))
    _lane(oracle_cli, LANES[5], extra=("--chunk-size", "131072", "--compression-level", "3"))


@pytest.mark.parametrize("case", FIXTURES, ids=[c.name for c in FIXTURES])
def test_device_logic_on_cpu_reproduces_reference_fixture(emu_cli, case):
    _fixture(emu_cli, case)


@pytest.mark.parametrize("l", LANES, ids=[l.name for l in LANES])
def test_device_logic_on_cpu_equals_reference_binary(emu_cli, l):
    _lane(emu_cli, l)


def test_device_logic_on_cpu_many_chunks(emu_cli):
    _lane(emu_cli, LANES[2], extra=("--chunk-size", "131072"))
    _lane(emu_cli, LANES[5], extra=("--chunk-size", "131072"))


@pytest.mark.gpu
@pytest.mark.parametrize("case", FIXTURES, ids=[c.name for c in FIXTURES])
def test_cuda_cli_reproduces_reference_fixture(product_cli, case):
    _fixture(product_cli, case)


@pytest.mark.gpu
@pytest.mark.parametrize("l", GPU_LANES, ids=[l.name for l in GPU_LANES])
def test_cuda_cli_equals_reference_binary(product_cli, l):
    _lane(product_cli, l)


@pytest.mark.gpu
def test_cuda_cli_many_chunks_and_host_gzip(product_cli):
    _lane(product_cli, LANES[2], extra=("--chunk-size", "131072"))
    _lane(product_cli, LANES[4], extra=("--host-gzip",))


def test_bad_headers_are_refused(oracle_cli, emu_cli):
    """A header barcode of another length than --barcode makes the reference panic (hamming_distance(..).unwrap(),
    src/lib.rs:1108-1112): the drop-in stops with an error instead of writing anything wrong."""
    l = RF.RfLane("bad", 50, 5, barcode="ACGTACGTAC", flags=["--barcode", "ACGTACG"])
    for binary in (oracle_cli, emu_cli):
        d = tempfile.mkdtemp(prefix="rf_")
        try:
            p1, p2 = RF.write_rf_lane(l, d)
            p = subprocess.run([binary] + RF.rf_args(l, p1, p2, os.path.join(d, "out")), capture_output=True, text=True)
            assert p.returncode != 0 and "header" in p.stderr
        finally:
            shutil.rmtree(d, ignore_errors=True)


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference binary not available")
def test_goldens_are_what_the_reference_binary_prints_now():
    _lane(REF_BIN, LANES[0])
    _fixture(REF_BIN, FIXTURES[1])


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference binary not available")
def test_reference_report_command_panics():
    """Why `report` (the other half of §8 f-4) is not built: the reference's own binary cannot merge the sample_stats
    files it writes."""
    d = tempfile.mkdtemp(prefix="rf_")
    try:
        case = FIXTURES[0]
        lanes.run_binary(REF_BIN, ["reformat"] + RF.resolve(case.args, DATA) + ["-o", os.path.join(d, "rf")])
        stats = os.path.join(d, "rf", "sample1.mgikit.sample_stats")
        assert os.path.exists(stats)
        p = subprocess.run([REF_BIN, "report", "--qc-report", stats, "-o", os.path.join(d, "o")], capture_output=True, text=True)
        assert p.returncode != 0 and "index out of bounds" in p.stderr
        assert not os.path.isdir(os.path.join(d, "o")) or not os.listdir(os.path.join(d, "o"))
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ---------------------------------------------------------------- through the C ABI, chunk level
def _abi_chunks(n, umi):
    from mgikit_b200 import synth
    r1, r2 = synth.reformat_block(synth.CONFIG2, 0, n)
    spec = synth.reformat_run_spec(synth.CONFIG2, umi_length=umi, max_chunk_bytes=max(r1.size, r2.size) + 4096, max_chunk_records=n + 64)
    return spec, [(r2.tobytes(), r1.tobytes())]


@pytest.mark.parametrize("umi", [0, 8, 20])
def test_device_logic_on_cpu_equals_oracle_through_the_abi(oracle_lib, emu_lib, umi):
    from util import run_both
    spec, chunks = _abi_chunks(3000, umi)
    run_both(emu_lib, "emu_", oracle_lib, "orc_", spec, chunks, f"reformat umi {umi}")


@pytest.mark.gpu
@pytest.mark.parametrize("umi", [0, 8, 20])
def test_cuda_equals_oracle_through_the_abi(oracle_lib, gpu_lib, umi):
    from util import run_both
    spec, chunks = _abi_chunks(50000, umi)
    run_both(gpu_lib, "mgk_", oracle_lib, "orc_", spec, chunks * 2, f"reformat umi {umi}")


@pytest.mark.gpu
def test_cuda_one_million_pairs_properties(gpu_lib):
    """Full chunk size: every record comes back once, in order, with the header rewritten and the UMI moved."""
    import numpy as np
    from mgikit_b200 import synth
    from util import make_hotpath
    n, umi = 1 << 20, 8
    r1, r2 = synth.reformat_block(synth.CONFIG2, 0, n)
    spec = synth.reformat_run_spec(synth.CONFIG2, umi_length=umi, max_chunk_bytes=max(r1.size, r2.size) + 4096, max_chunk_records=n + 64)
    hp = make_hotpath(gpu_lib, "mgk_", spec)
    try:
        res = hp.process(r2.tobytes(), r1.tobytes())
        assert res.n_records == n and res.consumed_r2 == r2.size and res.consumed_r1 == r1.size
        assert int(res.seg_len_r2[1]) == 0 and int(res.seg_len_r2[2]) == 0
        o2 = np.frombuffer(res.segment(2, 0), np.uint8)
        o1 = np.frombuffer(res.segment(1, 0), np.uint8)
        assert int((o2 == 10).sum()) == 4 * n and int((o1 == 10).sum()) == 4 * n
        # the paired read's body is untouched: total bytes = input - old headers + new headers
        nl2 = np.nonzero(o2 == 10)[0]
        seq_len = nl2[1::4] - nl2[0::4] - 1
        in_seq = (r2.size // n) - 0  # fixed-size records
        assert (seq_len == seq_len[0]).all() and int(seq_len[0]) == synth.CONFIG2.r2_insert + 20 - umi
        stats, mism = hp.counters()
        assert int(mism[0].sum()) == n and int(mism[0][1]) + int(mism[0][2]) == n   # 0 or 1 mismatch by construction
        assert int(mism[0][2]) > 0
        del in_seq
    finally:
        hp.close()
