"""The pump's failure paths (host side, oracle-bound CLI): damaged or mismatched inputs end the command with the
reference's exit status (101, like a Rust panic) and a message -- never a hang, never partial success."""
import gzip
import os
import shutil
import subprocess
import tempfile

import pytest

import lanes
from mgikit_b200 import synth


@pytest.fixture(scope="module")
def lane_dir():
    d = tempfile.mkdtemp(prefix="mgk_err_")
    r1, r2, sheet = lanes.write_lane(synth.CONFIG2, 20000, d)
    yield d, r1, r2, sheet
    shutil.rmtree(d, ignore_errors=True)


def _run(cli, lane_dir, r1, r2, extra=()):
    d, _, _, sheet = lane_dir
    a = lanes.demux_args(synth.CONFIG2, r1, r2, sheet, os.path.join(d, "out"), list(extra))
    p = subprocess.run([cli] + a, capture_output=True, text=True, timeout=180)
    return p.returncode, p.stderr


def test_truncated_and_damaged_inputs_are_errors(oracle_cli, lane_dir):
    d, r1, r2, _ = lane_dir
    assert _run(oracle_cli, lane_dir, r1, r2)[0] == 0
    raw = open(r2, "rb").read()
    cut = os.path.join(d, "cut_read_2.fq.gz")
    open(cut, "wb").write(raw[:len(raw) // 2])
    rc, err = _run(oracle_cli, lane_dir, r1, cut)
    assert rc == 101 and "unexpected end of file" in err
    bad = bytearray(raw)
    bad[len(bad) // 3] ^= 0x10
    flipped = os.path.join(d, "flip_read_2.fq.gz")
    open(flipped, "wb").write(bytes(bad))
    rc, err = _run(oracle_cli, lane_dir, r1, flipped)
    assert rc == 101 and "corrupt gzip input" in err


def test_reads_of_different_length_are_an_error(oracle_cli, lane_dir):
    d, r1, r2, _ = lane_dir
    text = gzip.open(r1).read()
    short = os.path.join(d, "short_read_1.fq.gz")
    open(short, "wb").write(gzip.compress(text[:len(text) // 2 // 336 * 336], 1))
    rc, err = _run(oracle_cli, lane_dir, short, r2, ["--chunk-size", "1048576"])
    assert rc == 101 and "Something wrong in the input files" in err      # src/lib.rs:1490-1494


def test_reformat_refusals(oracle_cli):
    """`reformat`'s own argument checks (src/lib.rs:2352-2361, src/formater.rs:23-25) and what the pump says when the
    records of a file outgrow the chunk they were sized for."""
    import reformat_cases as RF
    d = tempfile.mkdtemp(prefix="mgk_rferr_")
    try:
        l = RF.RfLane("e", 300, 9)
        p1, p2 = RF.write_rf_lane(l, d)

        def run(args):
            p = subprocess.run([oracle_cli] + args, capture_output=True, text=True, timeout=120)
            return p.returncode, p.stderr
        base = RF.rf_args(l, p1, p2, os.path.join(d, "out"))
        assert run(base)[0] == 0
        rc, err = run(base + ["--sample-index", "0"])
        assert rc == 101 and "needs to be greater than 0" in err
        # a file name without a third '_' part and no --sample-label
        q1, q2 = os.path.join(d, "a_1.fq.gz"), os.path.join(d, "a_2.fq.gz")
        shutil.copy(p1, q1)
        shutil.copy(p2, q2)
        rc, err = run(RF.rf_args(l, q1, q2, os.path.join(d, "out2")))
        assert rc == 101 and "Sample label needs to be passed" in err
        assert run(RF.rf_args(l, q1, q2, os.path.join(d, "out2")) + ["--sample-label", "x"])[0] == 0
        assert os.path.exists(os.path.join(d, "out2", "x_S1_L01_R1_001.fastq.gz"))
        rc, err = run(base + ["--compression-buffer-size", "999999999"])
        assert rc == 101 and "Compression buffer size" in err
        rc, err = run(base + ["--umi-length", "500"])
        assert rc == 101 and "UMI is longer" in err
        # records that grow beyond the first one's size by more than the chunk's slack
        odd = [x for x in RF.synthetic_lanes() if x.name == "rf-odd-headers"][0]
        o1, o2 = RF.write_rf_lane(odd, os.path.join(d, "odd"))
        rc, err = run(RF.rf_args(odd, o1, o2, os.path.join(d, "out3")) + ["--chunk-size", "65536"])
        assert rc == 101 and "raise --chunk-size" in err
    finally:
        shutil.rmtree(d, ignore_errors=True)
