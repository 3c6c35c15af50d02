"""The reference's own black-box test matrix for `demultiplex`, restated.

Source: /root/reference/tests/integration_test.rs — testing_demultiplex (:162-326),
testing_demultiplex_not_mgi_input (:329-434), testing_demultiplex_large (:968-1066),
testing_demultiplex_large_se (:1069-1172), testing_single_end (:1175-1337),
testing_mgi_full_header (:1340-1457).  Paths are relative to a `testing_data` root,
so the same cases run against /root/reference/testing_data (here) and against the
committed copy under tests/golden/ (GPU box).
"""
from __future__ import annotations

import dataclasses
import gzip
import hashlib
import os
import subprocess


@dataclasses.dataclass
class Case:
    name: str
    args: list            # CLI args after `demultiplex`, without -o
    expected: str         # directory under testing_data/expected
    # how to compare: list of (expected_file, produced_file or None to look up by suffix)
    skip_expected: tuple = ()       # substrings of expected file names to ignore
    rename: tuple = ()              # (old, new) applied to produced .gz names
    stats_override: str = ""        # compare produced sample_stats against this expected file instead
    by_suffix: bool = False         # report names carry a timestamp: locate by suffix
    extra_files: int = 0            # produced - expected file count
    check_count: bool = True


def _ds(n: int) -> str:
    return f"ds0{n}"


def demultiplex_cases() -> list:
    cases = []
    for ds in range(1, 15):
        disable_illumina = ds == 5
        ds_in = {6: 1, 9: 8, 7: 1, 4: 3, 5: 1, 11: 1, 12: 2, 14: 2}.get(ds, ds)
        ds_ex = {6: 1, 14: 12}.get(ds, ds)
        ds_fc = {10: 1, 11: 1, 12: 1, 13: 2, 14: 2}.get(ds_in, ds_in)
        r1 = f"input/{_ds(ds_in)}/L01/FC0{ds_fc}_L01_read_1.fq.gz"
        r2 = f"input/{_ds(ds_in)}/L01/FC0{ds_fc}_L01_read_2.fq.gz"
        in_dir = f"input/{_ds(ds_in)}/L01/" if ds == 6 else ""
        ext = "csv" if ds == 1 else "tsv"
        sheet = f"expected/{_ds(ds_ex)}/sample_sheet_expected.{ext}"
        instrument, run = ("instrument_3", "20230727") if ds in (3, 4) else ("instrument_1", "20231212")
        for m in range(5):
            args = ["-f", r1, "-r", r2, "-i", in_dir, "-s", sheet, "--lane", "L01", "--run", run,
                    "--instrument", instrument, "--writing-buffer-size", "131072", "-m", str(m), "--force",
                    "-t", "1"]
            if ds in (2, 5):
                args.append("--comprehensive-scan")
            args.append("--validate")
            if disable_illumina:
                args.append("--disable-illumina")
            if ds == 9:
                args += ["--template", "i78:--8"]
            if ds >= 11:
                args.append("--per-index-error")
            cases.append(Case(f"demux-{_ds(ds)}-m{m}", args, f"{_ds(ds_ex)}/{_ds(ds_ex)}-{m}"))
            if ds in (7, 8, 9, 10):
                break
            if ds > 10 and m == 2:
                break
    return cases


def not_mgi_cases() -> list:
    cases = []
    for m in range(3):
        args = ["-f", "input/ds015/L01/FC02_L01_read_1.fq.gz", "-r", "input/ds015/L01/FC02_L01_read_2.fq.gz",
                "-s", "expected/ds015/sample_sheet_expected.tsv", "--lane", "L01", "--writing-buffer-size",
                "131072", "-m", str(m), "--force", "--disable-illumina", "--not-mgi", "--validate",
                "--per-index-error"]
        cases.append(Case(f"notmgi-ds015-m{m}", args, f"ds015/ds015-{m}", by_suffix=True, extra_files=1))
    return cases


def large_cases() -> list:
    base = ["-s", "expected/ds01/sample_sheet_expected.tsv", "--lane", "L01", "--run", "20231212", "--instrument",
            "instrument_1", "--writing-buffer-size", "131072", "-m", "0", "--force", "--threads", "1", "--validate"]
    r1, r2 = "input/large_ds/ZFC01_L01_read_1.fq.gz", "input/large_ds/ZFC01_L01_read_2.fq.gz"
    pe = Case("large-pe", base + ["-f", r1, "-r", r2], "large_ds", skip_expected=("se-FC01.L01.mgikit",),
              check_count=False)
    se = Case("large-se", base + ["-f", r2], "large_ds", skip_expected=("se-FC01.L01.mgikit", "_R1_001"),
              rename=(("R2", "R1"),), stats_override="large_ds/se-FC01.L01.mgikit.sample_stats", check_count=False)
    return [pe, se]


def single_end_cases() -> list:
    cases = []
    for ds in range(1, 15):
        disable_illumina = ds == 5
        ds_in = {6: 1, 9: 8, 7: 1, 4: 3, 5: 1, 11: 1, 12: 2, 14: 2}.get(ds, ds)
        ds_ex = {6: 1, 14: 12}.get(ds, ds)
        ds_fc = {10: 1, 11: 1, 12: 1, 13: 2, 14: 2}.get(ds_in, ds_in)
        r2 = f"input/{_ds(ds_in)}/L01/FC0{ds_fc}_L01_read_2.fq.gz"
        ext = "csv" if ds == 1 else "tsv"
        sheet = f"expected/{_ds(ds_ex)}/sample_sheet_expected.{ext}"
        instrument, run = ("instrument_3", "20230727") if ds in (3, 4) else ("instrument_1", "20231212")
        for m in range(2):
            if m == 1 and ds != 2:
                continue
            args = ["-f", r2, "-s", sheet, "--lane", "L01", "--run", run, "--instrument", instrument,
                    "--writing-buffer-size", "131072", "-m", str(m), "--force", "-t", "1"]
            if ds in (2, 5):
                args.append("--comprehensive-scan")
            if disable_illumina:
                args.append("--disable-illumina")
            if ds == 9:
                args += ["--template", "i78:--8"]
            args.append("--validate")
            if ds >= 11:
                args.append("--per-index-error")
            cases.append(Case(f"se-{_ds(ds)}-m{m}", args, f"se-test/{_ds(ds_ex)}/{_ds(ds_ex)}-{m}",
                              check_count=(ds == 1)))
            if ds in (7, 8, 9, 10):
                break
    return cases


def full_header_cases() -> list:
    cases = []
    for ds in (1, 2, 8, 11):
        ds_in = 1 if ds == 11 else ds
        r1 = f"input/{_ds(ds_in)}/L01/FC0{ds_in}_L01_read_1.fq.gz"
        r2 = f"input/{_ds(ds_in)}/L01/FC0{ds_in}_L01_read_2.fq.gz"
        args = ["-f", r1, "-r", r2, "-s", f"expected/{_ds(ds)}/sample_sheet_expected.tsv", "--lane", "L01",
                "--writing-buffer-size", "131072", "-m", "0", "--force", "-t", "1", "--comprehensive-scan",
                "--disable-illumina", "--mgi-full-header"]
        if ds >= 11:
            args.append("--per-index-error")
        args.append("--validate")
        cases.append(Case(f"fullhdr-{_ds(ds)}", args, f"full_mgi_header-{ds}"))
    return cases


def all_cases() -> list:
    return demultiplex_cases() + not_mgi_cases() + large_cases() + single_end_cases() + full_header_cases()


# --------------------------------------------------------------------------- comparators
def gunzip_bytes(path: str) -> bytes:
    """MultiGzDecoder::read_to_string (integration_test.rs:44-53): all members, concatenated."""
    with open(path, "rb") as f:
        raw = f.read()
    if not raw:
        return b""
    return gzip.decompress(raw)


PATH_FLAGS = ("-f", "-r", "-i", "-s")


def resolve_args(case: Case, root: str) -> list:
    out = []
    it = iter(case.args)
    for a in it:
        out.append(a)
        if a in PATH_FLAGS:
            v = next(it)
            out.append(os.path.join(root, v) if v else v)
    return out


def run_case(binary: str, case: Case, root: str, out_dir: str, extra_args=()) -> subprocess.CompletedProcess:
    cmd = [binary, "demultiplex"] + resolve_args(case, root) + ["-o", out_dir] + list(extra_args)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


def check_stdout_shape(stdout: str) -> None:
    # integration_test.rs:141-143 / :277-279 index lines[1].split(" ")[1]
    lines = stdout.split("\n")
    assert len(lines) > 1 and len(lines[1].split(" ")) > 1


def expected_listing(case: Case, root: str) -> dict:
    """expected file name -> ('gz', sha256 of decompressed) | ('raw', sha256 of bytes)."""
    d = os.path.join(root, "expected", case.expected)
    out = {}
    for name in sorted(os.listdir(d)):
        p = os.path.join(d, name)
        if not os.path.isfile(p):
            continue
        if name.endswith(".gz"):
            out[name] = ("gz", hashlib.sha256(gunzip_bytes(p)).hexdigest())
        else:
            with open(p, "rb") as f:
                out[name] = ("raw", hashlib.sha256(f.read()).hexdigest())
    return out


def compare_outputs(case: Case, listing: dict, out_dir: str, stats_override_digest: str = "") -> None:
    """The comparison each reference test does, on digests."""
    produced = sorted(os.listdir(out_dir))
    for name, (kind, digest) in listing.items():
        if any(s in name for s in case.skip_expected):
            continue
        if kind == "gz":
            pname = name
            for old, new in case.rename:
                pname = pname.replace(old, new)
            got = hashlib.sha256(gunzip_bytes(os.path.join(out_dir, pname))).hexdigest()
            assert got == digest, f"{case.name}: decompressed {pname} differs from expected {name}"
        else:
            pname = name
            if case.by_suffix:
                # names carry a timestamp flowcell (src/run_manager.rs:263): locate by substring, :386-399
                cands = [p for p in produced if p.endswith(name)]
                assert len(cands) == 1, f"{case.name}: cannot locate {name} in {produced}"
                pname = cands[0]
            with open(os.path.join(out_dir, pname), "rb") as f:
                got = hashlib.sha256(f.read()).hexdigest()
            want = digest
            if case.stats_override and "sample_stats" in name:
                want = stats_override_digest
            assert got == want, f"{case.name}: report {pname} differs from expected"
    if case.check_count:
        n_exp = len(listing)
        assert len(produced) == n_exp + case.extra_files, f"{case.name}: {len(produced)} files, expected {n_exp + case.extra_files}: {produced}"
