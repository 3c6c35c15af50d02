"""`mgikit template`: barcode-template detection, the pre-step of BASELINE
config 4.  The reference's own test (tests/integration_test.rs:73-160) compares
<prefix>_template.tsv with expected/ds0N/sample_sheet_expected.tsv byte for byte; the
synthetic lanes are held to what the reference's binary answered (template_reference.json,
made by golden/make_golden.py).  The command is host code (no kernel): it runs here."""
import hashlib
import json
import os
import shutil
import tempfile

import pytest

import lanes
from golden.make_golden import run_template, synthetic_template_variants, synthetic_variants, template_cases

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "golden", "testing_data")
with open(os.path.join(HERE, "golden", "template_reference.json")) as f:
    TPL_GOLD = json.load(f)
with open(os.path.join(HERE, "golden", "synthetic_reference.json")) as f:
    SYN_GOLD = json.load(f)
REF_BIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "mgikit")


@pytest.mark.parametrize("case", template_cases(), ids=[c[0] for c in template_cases()])
def test_template_reproduces_reference_fixture(product_cli, case):
    name, r1, r2, sheet, extra, expected = case
    d = tempfile.mkdtemp(prefix="mgk_tpl_")
    try:
        args = ["template", "-f", os.path.join(DATA, r1), "-r", os.path.join(DATA, r2), "-s", os.path.join(DATA, sheet),
                "-o", os.path.join(d, "out", "sample_sheet")] + extra
        p = lanes.run_binary(product_cli, args)
        lines = p.stdout.split("\n")
        assert len(lines) > 1 and len(lines[1].split(" ")) > 1      # integration_test.rs:141-143 indexes these
        with open(os.path.join(d, "out", "sample_sheet_template.tsv"), "rb") as f, open(os.path.join(DATA, expected), "rb") as g:
            assert f.read() == g.read(), name
        assert os.path.exists(os.path.join(d, "out", "sample_sheet_details.tsv"))
    finally:
        shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("name,lane,n,flags", synthetic_template_variants(), ids=[v[0] for v in synthetic_template_variants()])
def test_template_equals_reference_binary_on_synthetic_lanes(product_cli, name, lane, n, flags):
    d = tempfile.mkdtemp(prefix="mgk_tpl_")
    try:
        assert run_template(product_cli, lane, n, flags, d) == TPL_GOLD[name]["template_tsv"], name
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_single_end_without_barcode_length_is_refused(product_cli):
    """src/lib.rs:2025-2045: the read-length difference of a single-end lane is the whole read -> the reference panics."""
    import subprocess
    from mgikit_b200 import synth
    d = tempfile.mkdtemp(prefix="mgk_tpl_")
    try:
        r1, _, sheet = lanes.write_lane(synth.CONFIG4, 200, d)
        p = subprocess.run([product_cli, "template", "-f", r1, "-s", sheet, "-o", os.path.join(d, "t", "x")], capture_output=True, text=True)
        assert p.returncode != 0 and "prvide barcode length" in p.stderr
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _config4_through_template(binary, template_binary):
    """BASELINE config 4 as stated: `template` on the first 10k reads, then demultiplex with the sheet it wrote
    (no --template flag) == what the reference wrote for the lane."""
    name, lane, extra, n = [v for v in synthetic_variants() if v[0] == "cfg4"][0]
    d = tempfile.mkdtemp(prefix="mgk_cfg4_")
    try:
        r1, r2, sheet = lanes.write_lane(lane, n, d)
        lanes.run_binary(template_binary, ["template", "-f", r1, "-s", sheet, "--testing-reads", "10000", "--barcode-length", "8",
                                           "-o", os.path.join(d, "tpl", "sheet")])
        out = os.path.join(d, "out")
        args = lanes.demux_args(lane, r1, r2, os.path.join(d, "tpl", "sheet_template.tsv"), out)
        k = args.index("--template")
        del args[k:k + 2]
        lanes.run_binary(binary, args)
        got, want = lanes.digest_dir(out), SYN_GOLD[name]["files"]
        assert sorted(got) == sorted(want)
        diff = [f for f in want if got[f] != want[f]]
        assert not diff, f"cfg4 through the detected template: {len(diff)} files differ, e.g. {diff[:4]}"
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_config4_template_then_demultiplex_oracle(oracle_cli, product_cli):
    _config4_through_template(oracle_cli, product_cli)


@pytest.mark.gpu
def test_config4_template_then_demultiplex_cuda(product_cli):
    _config4_through_template(product_cli, product_cli)


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference binary not available")
def test_template_goldens_are_what_the_reference_binary_prints_now():
    name, lane, n, flags = synthetic_template_variants()[0]
    d = tempfile.mkdtemp(prefix="mgk_tpl_")
    try:
        assert run_template(REF_BIN, lane, n, flags, d) == TPL_GOLD[name]["template_tsv"]
    finally:
        shutil.rmtree(d, ignore_errors=True)
