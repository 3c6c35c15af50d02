#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference (run in the build container only).

 1. copies the INPUT fixtures of the reference's demultiplex tests
    (/root/reference/testing_data/input/*, sample sheets) so the GPU box, which
    has no /root/reference, can replay them;
 2. records, per case of tests/ref_matrix.py, the digest of every file under
    the reference's testing_data/expected (sha256 of the DECOMPRESSED text for
    .gz — the comparison the reference's tests make — and of the raw bytes for
    reports)  -> reference_matrix.json;
 3. runs the reference's own binary (oracle/_ref/mgikit, v2.2.0) on seeded
    synthetic lanes shaped like BASELINE.json configs 2-5 and records the same
    digests -> synthetic_reference.json.
 4. `reformat` (also alone: `make_golden.py reformat`): the inputs and expected
    digests of the reference's extras_test fixtures -> reformat_matrix.json, and
    the reference binary's answer on seeded one-sample lanes
    -> reformat_reference.json.
"""
import dataclasses
import hashlib
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import lanes  # noqa: E402
import ref_matrix as M  # noqa: E402
import reformat_cases as RF  # noqa: E402
from mgikit_b200 import synth  # noqa: E402

REF = "/root/reference/testing_data"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "mgikit")
DST = os.path.join(HERE, "testing_data")


def synthetic_variants():
    small = dataclasses.replace(synth.CONFIG2, frac_random=0.5, frac_n=0.1, n_samples=8)
    return [("cfg2", synth.CONFIG2, [], 20000), ("cfg3", synth.CONFIG3, [], 20000), ("cfg4", synth.CONFIG4, [], 20000),
            ("cfg5", synth.CONFIG5, [], 20000), ("cfg2-mgi", synth.CONFIG2, ["--disable-illumina"], 5000),
            ("cfg3-fullhdr", synth.CONFIG3, ["--disable-illumina", "--mgi-full-header"], 5000),
            ("cfg2-keep", synth.CONFIG2, ["--keep-barcode"], 5000),
            ("cfg3-perindex", synth.CONFIG3, ["--per-index-error"], 5000),
            ("cfg2-report1", synth.CONFIG2, ["--report-level", "1"], 5000),
            ("cfg2-noisy", small, ["--report-limit", "7"], 5000)] + odd_lanes()


def odd_lanes():
    """Lanes nobody would sequence, to pin the corners against the reference's binary: template orders, short and
    barcode-only reads, m = 0..3 (ambiguity), single-end dual index, flags in combination."""
    L = synth.LaneSpec
    noisy = dict(frac_one_mm=0.2, frac_two_mm=0.1, frac_random=0.1, frac_n=0.05)
    return [("odd-i5-i7-umi-m2", L("odd1", 12, "i58:i78:um8", r1_len=75, r2_insert=60, mismatches=2, seed=61001, **noisy), [], 3000),
            ("odd-umi16-i7", L("odd2", 10, "um16:i78", r1_len=50, r2_insert=40, seed=61002, **noisy), [], 3000),
            ("odd-single-end-dual-rc", L("odd3", 16, "i78:i58", r1_len=0, r2_insert=100, i5_rc=True, seed=61003, **noisy), [], 3000),
            ("odd-m0", L("odd4", 6, "i712:i512", r1_len=30, r2_insert=30, mismatches=0, seed=61004, **noisy), [], 3000),
            ("odd-m3-per-index", L("odd5", 8, "i710:i510", r1_len=40, r2_insert=35, mismatches=3, seed=61005, **noisy),
             ["--per-index-error"], 3000),
            ("odd-full-header-umi", L("odd6", 9, "i78:um8:i58", r1_len=60, r2_insert=45, mismatches=1, seed=61006, **noisy),
             ["--disable-illumina", "--mgi-full-header"], 3000),
            ("odd-keep-mgi", L("odd7", 7, "i710:i510", r1_len=33, r2_insert=21, seed=61007, **noisy),
             ["--disable-illumina", "--keep-barcode"], 3000),
            ("odd-barcode-only-read", L("odd8", 11, "i78:i58", r1_len=90, r2_insert=0, seed=61008, **noisy), [], 3000),
            ("odd-report0-validate", L("odd9", 5, "um4:i76:i56", r1_len=25, r2_insert=17, mismatches=2, seed=61009, **noisy),
             ["--report-level", "0", "--validate"], 3000),
            ("odd-long-reads", L("odd10", 14, "i710:i510", r1_len=301, r2_insert=250, seed=61010, **noisy), ["--comprehensive-scan"], 2000)]


def template_cases():
    """testing_template (tests/integration_test.rs:73-160): (name, fixture set, sample sheet, extra flags, expected file)."""
    out = []
    for ds in (1, 2, 3, 7):
        ds_in = 1 if ds == 7 else ds
        sheet = "input/ds01/sample_sheet-ds07.tsv" if ds == 7 else f"input/ds0{ds_in}/sample_sheet.tsv"
        extra = ["--barcode-length", "20" if ds == 3 else "0"] + ([] if ds in (2, 3) else ["--popular-template"])
        out.append((f"template-ds0{ds}", f"input/ds0{ds_in}/L01/FC0{ds_in}_L01_read_1.fq.gz", f"input/ds0{ds_in}/L01/FC0{ds_in}_L01_read_2.fq.gz",
                    sheet, extra, f"expected/ds0{ds}/sample_sheet_expected.tsv"))
    return out


def synthetic_template_variants():
    """(name, lane, reads, flags of `mgikit template`): the reference binary's own answer is recorded."""
    return [("tpl-cfg4", synth.CONFIG4, 20000, ["--testing-reads", "10000", "--barcode-length", "8"]),
            ("tpl-cfg2", synth.CONFIG2, 6000, []),
            ("tpl-cfg3-popular", synth.CONFIG3, 6000, ["--popular-template", "--testing-reads", "3000"]),
            ("tpl-cfg3-no-umi", synth.CONFIG3, 6000, ["--no-umi", "--barcode-length", "28"])]


def run_template(binary, lane, n, flags, d):
    r1, r2, sheet = lanes.write_lane(lane, n, d)
    args = ["template", "-f", r1] + (["-r", r2] if r2 else []) + ["-s", sheet, "-o", os.path.join(d, "tpl", "sheet")] + list(flags)
    lanes.run_binary(binary, args)
    with open(os.path.join(d, "tpl", "sheet_template.tsv"), "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def copy_fixture(rel):
    src = os.path.join(REF, rel)
    if os.path.isdir(src):
        for name in os.listdir(src):
            if os.path.isfile(os.path.join(src, name)):
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                shutil.copy(os.path.join(src, name), os.path.join(DST, rel, name))
    else:
        os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
        shutil.copy(src, os.path.join(DST, rel))


def reformat_goldens():
    matrix = {}
    for c in RF.fixture_cases():
        it = iter(c.args)
        for a in it:
            if a in RF.PATH_FLAGS:
                copy_fixture(next(it))
        matrix[c.name] = {"files": lanes.digest_dir(os.path.join(REF, "expected", c.expected))}
    with open(os.path.join(HERE, "reformat_matrix.json"), "w") as f:
        json.dump(matrix, f, indent=0, sort_keys=True)
    syn = {}
    for l in RF.synthetic_lanes():
        d = tempfile.mkdtemp(prefix="golden_")
        p1, p2 = RF.write_rf_lane(l, d)
        out = os.path.join(d, "out")
        lanes.run_binary(REF_BIN, RF.rf_args(l, p1, p2, out))
        t1, t2 = RF.lane_text(l)
        syn[l.name] = {"inputs": [hashlib.sha256(t).hexdigest() for t in (t1, t2) if t is not None], "files": lanes.digest_dir(out)}
        shutil.rmtree(d)
    with open(os.path.join(HERE, "reformat_reference.json"), "w") as f:
        json.dump(syn, f, indent=0, sort_keys=True)
    print(f"{len(matrix)} reformat fixture cases, {len(syn)} synthetic reformat lanes")


def main():
    if sys.argv[1:] == ["reformat"]:
        reformat_goldens()
        return
    cases = M.all_cases()
    needed = set()
    for _, r1, r2, sheet, _, expected in template_cases():
        needed.update([r1, r2, sheet, expected])
    for c in cases:
        it = iter(c.args)
        for a in it:
            if a in M.PATH_FLAGS:
                v = next(it)
                if v:
                    needed.add(v)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in sorted(needed):
        src = os.path.join(REF, rel)
        if os.path.isdir(src):
            for name in os.listdir(src):
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                shutil.copy(os.path.join(src, name), os.path.join(DST, rel, name))
        else:
            os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
            shutil.copy(src, os.path.join(DST, rel))
    matrix = {}
    for c in cases:
        entry = {"files": M.expected_listing(c, REF)}
        if c.stats_override:
            with open(os.path.join(REF, "expected", c.stats_override), "rb") as f:
                entry["stats_override"] = hashlib.sha256(f.read()).hexdigest()
        matrix[c.name] = entry
    with open(os.path.join(HERE, "reference_matrix.json"), "w") as f:
        json.dump(matrix, f, indent=0, sort_keys=True)

    syn = {}
    for name, lane, extra, n in synthetic_variants():
        d = tempfile.mkdtemp(prefix="golden_")
        r1, r2, sheet = lanes.write_lane(lane, n, d)
        out = os.path.join(d, "out")
        lanes.run_binary(REF_BIN, lanes.demux_args(lane, r1, r2, sheet, out, extra))
        inputs = {}
        for p in (r1, r2):
            if p:
                import gzip
                with gzip.open(p, "rb") as f:
                    inputs[os.path.basename(p)] = hashlib.sha256(f.read()).hexdigest()
        syn[name] = {"n_pairs": n, "extra": extra, "inputs": inputs, "files": lanes.digest_dir(out)}
        shutil.rmtree(d)
    with open(os.path.join(HERE, "synthetic_reference.json"), "w") as f:
        json.dump(syn, f, indent=0, sort_keys=True)
    tpl = {}
    for name, lane, n, flags in synthetic_template_variants():
        d = tempfile.mkdtemp(prefix="golden_")
        tpl[name] = {"n_pairs": n, "flags": flags, "template_tsv": run_template(REF_BIN, lane, n, flags, d)}
        shutil.rmtree(d)
    with open(os.path.join(HERE, "template_reference.json"), "w") as f:
        json.dump(tpl, f, indent=0, sort_keys=True)
    print(f"{len(matrix)} fixture cases, {len(syn)} synthetic lanes, {len(tpl)} template lanes")
    reformat_goldens()


if __name__ == "__main__":
    main()
