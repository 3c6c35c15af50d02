"""The `reformat` cases: the reference's own (tests/integration_test.rs:437-599 testing_reformat, :601-795
testing_reformat_raw, fixtures under testing_data/{input,expected}/extras_test) and seeded synthetic lanes of one
sample whose answer comes from the reference's binary (tests/golden/reformat_reference.json)."""
from __future__ import annotations

import gzip
import os
import random
from dataclasses import dataclass, field

IN = "input/extras_test"


@dataclass
class RefCase:
    name: str
    args: list          # after `reformat`; paths relative to testing_data, -o appended by the runner
    expected: str       # directory under testing_data/expected
    rename: tuple = ()  # (from, to) applied to the expected file names (ds12: --sample-label)


def _common(r1, r2, sample_index):
    return ["-f", r1, "-r", r2, "--lane", "L01", "--force", "--sample-index", str(sample_index)]


def fixture_cases() -> list:
    out = []
    # testing_reformat: headers that already end with " n:N:0:BARCODE"
    for it in range(1, 8):
        expected_itr, input_itr = (1, 1) if it == 4 else ((it, 1) if it in (5, 6, 7) else (it, it))
        fc, instrument, run = ("FC03", "instrument_3", "20230727") if it == 3 else ("FC01", "instrument_1", "20231212")
        a = _common(f"{IN}/{fc}_L01_sample{input_itr}_1.fq.gz", f"{IN}/{fc}_L01_sample{input_itr}_2.fq.gz", input_itr)
        if it == 2:
            a += ["--umi-length", "8"]
        if it == 3:
            a += ["--umi-length", "20"]
        if it in (1, 2, 3, 5):
            a += ["--run", run, "--instrument", instrument]
        else:
            a += ["--info-file", f"{IN}/BioInfo.csv"]
        if it == 5:
            a += ["--disable-illumina"]
        if it == 6:
            a += ["--report-level", "0"]
        if it == 7:
            a += ["--barcode", "TCCGTTGAAA"]
        out.append(RefCase(f"sb-{it}", a, f"extras_test/ds{expected_itr}"))
    # testing_reformat_raw: MGI raw headers ("/n")
    for it in range(1, 13):
        if it == 8:
            continue
        if it == 4:
            expected_itr, input_itr = 1, 1
        elif it in (5, 6, 7, 12):
            expected_itr, input_itr = it, 1
        elif 8 < it < 12:
            expected_itr, input_itr = it, it - 8
        else:
            expected_itr, input_itr = it, it
        fc, instrument, run = ("FC03", "instrument_3", "20230727") if it in (3, 11) else ("FC01", "instrument_1", "20231212")
        a = _common(f"{IN}/raw/{fc}_L01_sample{input_itr}_1.fq.gz", f"{IN}/raw/{fc}_L01_sample{input_itr}_2.fq.gz", input_itr)
        if it in (2, 10):
            a += ["--umi-length", "8"]
        if it in (3, 11):
            a += ["--umi-length", "20"]
        if it in (1, 2, 3, 5, 11, 12):
            a += ["--run", run, "--instrument", instrument]
        else:
            a += ["--info-file", f"{IN}/BioInfo.csv"]
        if it == 5:
            a += ["--disable-illumina"]
        if it == 6:
            a += ["--report-level", "0"]
        if it == 7:
            a += ["--barcode", "TCCGTTGAAA"]
        if it in (1, 2, 4, 6):
            a += ["--barcode", "TCCGTTGAAT"]
        elif it == 3:
            a += ["--barcode", "AAAAAAAA"]
        if it == 12:
            a += ["--sample-label", "test-sample"]
        out.append(RefCase(f"raw-{it}", a, f"extras_test/ds{expected_itr}{'-raw' if it == 7 else ''}",
                           ("sample1", "test-sample") if it == 12 else ()))
    return out


PATH_FLAGS = ("-f", "-r", "--info-file")


def resolve(args, root):
    out, it = [], iter(args)
    for a in it:
        out.append(a)
        if a in PATH_FLAGS:
            out.append(os.path.join(root, next(it)))
    return out


# ---------------------------------------------------------------- synthetic lanes of one sample
@dataclass
class RfLane:
    name: str
    n: int
    seed: int
    style: str = "sb"          # "sb": "@... n:N:0:BARCODE", "raw": "@.../n"
    paired: bool = True
    r1_len: int = 100
    r2_len: int = 60
    barcode: str = "ACGTTGCAAC"   # what the header carries (with errors); "" with style raw
    frac_mm: float = 0.3          # reads whose header barcode differs from it
    flags: list = field(default_factory=list)
    instrument: str = "INS"
    run: str = "20240101"
    cpu_only: bool = False        # added after the last GPU session of the round: oracle and device logic on the CPU only


BASES = "ACGT"


def lane_text(l: RfLane):
    """(R1 text or None, barcode-read text) of the lane; plain loops: these lanes are a few thousand reads."""
    rng = random.Random(l.seed)
    r1, r2 = [], []
    for i in range(l.n):
        # column / row with and without leading zeros, read numbers from all zeros to no zero at all
        col = rng.choice([1, 7, 12, 45, 120, 999])
        row = rng.choice([1, 3, 10, 88, 100, 754])
        num = rng.choice([0, 5, 13, 1000, 99999, 1234567, 12345678, rng.randrange(10 ** 8)]) if i else 0
        digits = rng.choice([8, 8, 8, 10, 12]) if i % 97 == 5 else 8
        head = f"@V350099999L1C{col:03d}R{row:03d}{num:0{digits}d}"
        bc = list(l.barcode)
        if bc and rng.random() < l.frac_mm:
            for _ in range(rng.choice([1, 1, 2, 3, 5, 7])):
                p = rng.randrange(len(bc))
                bc[p] = rng.choice(BASES + "N")
        bc = "".join(bc)

        odd = i % 7 if (l.style == "odd" and i) else 0

        def rec(which, length):
            seq = "".join(rng.choice(BASES) if rng.random() > 0.01 else "N" for _ in range(length))
            qual = "".join(chr(33 + rng.choice([2, 11, 25, 30, 30, 37, 40])) for _ in range(length))
            tail = f" {which}:N:0:{bc}" if l.style in ("sb", "odd") else f"/{which}"
            if odd == 1:      # the blank three bytes before the end: "sep_position == len - 3" without being raw
                tail = f" {which}"
            elif odd == 2:    # no blank at all in an otherwise blank-separated file
                tail = f"/{which}"
            elif odd == 3:    # more fields behind the barcode
                tail = f" {which}:N:0:{bc} extra:field {i}"
            elif odd == 4:    # a long tail (taken over byte by byte: the header outgrows the staging slot)
                tail = f" {which}:N:0:{bc}:" + "x" * 150
            elif odd == 5:    # two blanks: the first one counts
                tail = f" {which}:Y:18:{bc} {which}:N:0:{bc}"
            return f"{head}{tail}\n{seq}\n+\n{qual}\n"

        if l.paired:
            r1.append(rec(1, l.r1_len))
            r2.append(rec(2, l.r2_len))
        else:
            r2.append(rec(1, l.r2_len))
    return ("".join(r1).encode() if l.paired else None), "".join(r2).encode()


def write_rf_lane(l: RfLane, d: str):
    os.makedirs(d, exist_ok=True)
    t1, t2 = lane_text(l)
    p1 = os.path.join(d, "V350099999_L01_smp-A_1.fq.gz")
    p2 = os.path.join(d, "V350099999_L01_smp-A_2.fq.gz")
    if l.paired:
        with gzip.open(p1, "wb", compresslevel=1) as f:
            f.write(t1)
        with gzip.open(p2, "wb", compresslevel=1) as f:
            f.write(t2)
        return p1, p2
    with gzip.open(p1, "wb", compresslevel=1) as f:
        f.write(t2)
    return p1, None


def rf_args(l: RfLane, p1, p2, out):
    a = ["reformat", "-f", p1] + (["-r", p2] if p2 else [])
    return a + ["--lane", "L01", "--run", l.run, "--instrument", l.instrument, "-o", out, "--force"] + list(l.flags)


def synthetic_lanes() -> list:
    bc = "ACGTTGCAAC"
    return [
        RfLane("rf-sb-barcode", 4000, 71001, flags=["--barcode", bc]),
        RfLane("rf-sb-plain", 3000, 71002, flags=["--sample-index", "7"]),
        RfLane("rf-sb-umi8-barcode", 3000, 71003, flags=["--barcode", bc, "--umi-length", "8"]),
        RfLane("rf-raw-plain", 3000, 71004, style="raw", barcode=""),
        RfLane("rf-raw-barcode-umi", 3000, 71005, style="raw", barcode="", flags=["--barcode", bc, "--umi-length", "12"]),
        RfLane("rf-se-sb-barcode", 3000, 71006, paired=False, r2_len=75, flags=["--barcode", bc, "--umi-length", "6"]),
        RfLane("rf-se-raw", 2000, 71007, paired=False, style="raw", barcode="", r2_len=40, flags=["--sample-label", "other", "--validate"]),
        RfLane("rf-sb-mgi-names", 2000, 71008, flags=["--disable-illumina", "--barcode", bc]),
        RfLane("rf-sb-level1", 2000, 71009, flags=["--report-level", "1", "--barcode", bc]),
        RfLane("rf-sb-level0", 2000, 71010, flags=["--report-level", "0"]),
        RfLane("rf-raw-validate", 2000, 71011, style="raw", barcode="", flags=["--validate", "--barcode", bc]),
        RfLane("rf-sb-umi-all", 1500, 71012, r2_len=20, flags=["--umi-length", "20"]),
        RfLane("rf-sb-dual", 2000, 71013, barcode="ACGTTGCAAC+TTGACCAGTA", flags=["--barcode", "ACGTTGCAAC+TTGACCAGTA"]),
        RfLane("rf-raw-long", 1500, 71014, style="raw", barcode="", r1_len=251, r2_len=301, flags=["--barcode", "ACGTACGT"]),
        # rewritten headers longer than the 80-byte staging slots of k_emit: built at their destination
        RfLane("rf-sb-long-header", 1500, 71015, barcode="ACGTTGCAAC+TTGACCAGTA", instrument="NOVASEQ-LIKE-INSTRUMENT-0123456789", run="2024010112345678",
               flags=["--barcode", "ACGTTGCAAC+TTGACCAGTA", "--umi-length", "12"]),
        # headers nobody writes on purpose, one oddity per read (no --barcode: the reference panics on a header barcode of
        # another length); with --barcode only the oddities that keep the barcode's place
        RfLane("rf-odd-headers", 1400, 71017, style="odd", flags=["--umi-length", "5"], cpu_only=True),
        RfLane("rf-odd-headers-se", 1400, 71018, style="odd", paired=False, r2_len=44, flags=[], cpu_only=True),
        RfLane("rf-raw-long-header", 1500, 71016, style="raw", barcode="", instrument="NOVASEQ-LIKE-INSTRUMENT-0123456789", run="2024010112345678",
               flags=["--barcode", "ACGTTGCAACTTGACCAGTAACGT", "--umi-length", "3"]),
    ]
