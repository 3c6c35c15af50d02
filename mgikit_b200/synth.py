"""Seeded synthetic MGI lanes, block-streamed.

Block ``b`` of a lane is a pure function of ``(seed, b)``: read numbers continue
across blocks, so any block can be regenerated for a parity spot check and the
benchmarks never touch the disk.  Modelled on the reference's generator
(/root/reference/testing_data/generate_fastq/generate_fastq.py:73-97) but
vectorised with numpy; header ``@V350012345L1C###R###########/n`` has a fixed
width like real MGI data.
"""
from __future__ import annotations

import dataclasses
from typing import List, Tuple

import numpy as np

from .plan import RunSpec, Sample, reverse_complement

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
FLOWCELL = "V350012345"
L_POSITION = 1 + len(FLOWCELL)      # index of the last 'L' in the header line


@dataclasses.dataclass
class LaneSpec:
    name: str
    n_samples: int
    template: str                 # e.g. "i710:i510"
    r1_len: int = 150             # paired read length (0 => single-end lane)
    r2_insert: int = 150          # bases of the barcode read before the barcode
    mismatches: int = 1
    i5_rc: bool = False
    seed: int = 20001
    frac_one_mm: float = 0.10     # one substitution in i7
    frac_two_mm: float = 0.0      # one substitution in i7 and one in i5
    frac_random: float = 0.03     # random barcode
    frac_n: float = 0.001         # one N inside the barcode
    min_dist: int = 3

    @property
    def paired(self) -> bool:
        return self.r1_len > 0


# the configurations BASELINE.json names
CONFIG2 = LaneSpec("cfg2-96x-dual10-m1", 96, "i710:i510", seed=20001)
CONFIG3 = LaneSpec("cfg3-384x-dual10-umi8-m2-i5rc", 384, "i710:um8:i510", mismatches=2, i5_rc=True, seed=30001,
                   frac_one_mm=0.05, frac_two_mm=0.05)
CONFIG4 = LaneSpec("cfg4-se-96x-single8-m1", 96, "i78", r1_len=0, r2_insert=50, seed=40001)
CONFIG5 = LaneSpec("cfg5-384x-dual10-m1", 384, "i710:i510", seed=50001)


def _template_items(tmpl: str) -> List[Tuple[str, int]]:
    return [(it[:2], int(it[2:])) for it in tmpl.split(":")]


def make_indexes(n: int, length: int, min_dist: int, rng: np.random.Generator) -> List[str]:
    """n distinct ACGT strings with pairwise Hamming distance >= min_dist (greedy)."""
    chosen = np.zeros((0, length), dtype=np.uint8)
    while chosen.shape[0] < n:
        cand = rng.integers(0, 4, size=(4 * n, length), dtype=np.uint8)
        for c in cand:
            if chosen.shape[0] and (np.sum(chosen != c, axis=1) < min_dist).any():
                continue
            chosen = np.vstack([chosen, c[None, :]])
            if chosen.shape[0] == n:
                break
    return ["".join("ACGT"[v] for v in row) for row in chosen]


def make_samples(spec: LaneSpec) -> List[Sample]:
    rng = np.random.default_rng(spec.seed)
    items = dict(_template_items(spec.template))
    i7s = make_indexes(spec.n_samples, items["i7"], spec.min_dist, rng)
    i5s = make_indexes(spec.n_samples, items["i5"], spec.min_dist, rng) if "i5" in items else None
    return [Sample(f"S{k + 1:03d}", i7s[k], i5s[k] if i5s else ".") for k in range(spec.n_samples)]


def run_spec(spec: LaneSpec, **kw) -> RunSpec:
    return RunSpec(samples=make_samples(spec), template=spec.template, i5_rc=spec.i5_rc, paired=spec.paired,
                   mismatches=spec.mismatches, l_position=L_POSITION,
                   illumina_prefix=f"@INS:20240101:{FLOWCELL}:", **kw)


def sample_sheet_text(spec: LaneSpec) -> str:
    rows = ["sample_id\ti7\ti5\tjob_number"]
    for s in make_samples(spec):
        rows.append(f"{s.sample_id}\t{s.i7}\t{s.i5}\t.")
    return "\n".join(rows) + "\n"


def _digits(vals: np.ndarray, width: int) -> np.ndarray:
    out = np.empty((vals.shape[0], width), dtype=np.uint8)
    v = vals.astype(np.int64).copy()
    for k in range(width - 1, -1, -1):
        out[:, k] = 48 + (v % 10)
        v //= 10
    return out


def _seq_matrix(strings: List[str]) -> np.ndarray:
    return np.frombuffer("".join(strings).encode(), dtype=np.uint8).reshape(len(strings), -1).copy()


def make_block(spec: LaneSpec, block: int, n_pairs: int, first_read: int = None):
    """Returns (r1 bytes or None, r2 bytes) as uint8 arrays of whole records."""
    rng = np.random.default_rng([spec.seed, block, 7])
    samples = make_samples(spec)
    items = _template_items(spec.template)
    BL = sum(n for _, n in items)
    n = n_pairs
    start = first_read if first_read is not None else block * n_pairs
    idx = np.arange(start, start + n, dtype=np.int64)

    # header: @FLOWCELL L1 Cccc Rrrr nnnnnnnn /x \n   (32 bytes)
    hl = 1 + len(FLOWCELL) + 2 + 4 + 4 + 8 + 2 + 1
    hdr = np.empty((n, hl), dtype=np.uint8)
    fixed = ("@" + FLOWCELL + "L1C").encode()
    hdr[:, :len(fixed)] = np.frombuffer(fixed, dtype=np.uint8)
    p = len(fixed)
    hdr[:, p:p + 3] = _digits(1 + (idx % 120), 3)
    hdr[:, p + 3] = ord("R")
    hdr[:, p + 4:p + 7] = _digits(1 + ((idx // 120) % 90), 3)
    hdr[:, p + 7:p + 15] = _digits((idx + 1) % 100000000, 8)
    hdr[:, p + 15] = ord("/")
    hdr[:, p + 16] = ord("2") if spec.paired else ord("1")
    hdr[:, p + 17] = 10

    # barcodes
    sid = rng.integers(0, spec.n_samples, size=n)
    i7m = _seq_matrix([s.i7 for s in samples])
    has_i5 = any(k == "i5" for k, _ in items)
    if has_i5:
        i5_in_read = [reverse_complement(s.i5) if spec.i5_rc else s.i5 for s in samples]
        i5m = _seq_matrix(i5_in_read)
    bar = np.empty((n, BL), dtype=np.uint8)
    col = 0
    pos = {}
    for kind, ln in items:
        pos[kind] = (col, ln)
        if kind == "i7":
            bar[:, col:col + ln] = i7m[sid]
        elif kind == "i5":
            bar[:, col:col + ln] = i5m[sid]
        else:
            bar[:, col:col + ln] = BASES[rng.integers(0, 4, size=(n, ln))]
        col += ln
    u = rng.random(n)

    def substitute(rows, kind):
        c0, ln = pos[kind]
        at = c0 + rng.integers(0, ln, size=rows.size)
        old = bar[rows, at]
        new = BASES[rng.integers(0, 4, size=rows.size)]
        same = new == old
        new[same] = BASES[(np.searchsorted(BASES, old[same]) + 1) % 4]
        bar[rows, at] = new

    t1 = spec.frac_one_mm
    t2 = t1 + spec.frac_two_mm
    t3 = t2 + spec.frac_random
    t4 = t3 + spec.frac_n
    one = np.nonzero(u < t1)[0]
    substitute(one, "i7")
    two = np.nonzero((u >= t1) & (u < t2))[0]
    if two.size:
        substitute(two, "i7")
        substitute(two, "i5" if has_i5 else "i7")
    rnd = np.nonzero((u >= t2) & (u < t3))[0]
    bar[rnd] = BASES[rng.integers(0, 4, size=(rnd.size, BL))]
    nn = np.nonzero((u >= t3) & (u < t4))[0]
    bar[nn, rng.integers(0, BL, size=nn.size)] = ord("N")

    def record_block(header, seq_len, tail):
        total = seq_len + (tail.shape[1] if tail is not None else 0)
        rec_len = hl + total + 1 + 2 + total + 1
        out = np.empty((n, rec_len), dtype=np.uint8)
        out[:, :hl] = header
        q = hl
        out[:, q:q + seq_len] = BASES[rng.integers(0, 4, size=(n, seq_len))]
        if tail is not None:
            out[:, q + seq_len:q + total] = tail
        q += total
        out[:, q] = 10
        out[:, q + 1] = ord("+")
        out[:, q + 2] = 10
        q += 3
        out[:, q:q + total] = rng.integers(53, 74, size=(n, total), dtype=np.uint8)   # Phred 20..40
        out[:, q + total] = 10
        return out.reshape(-1)

    r2 = record_block(hdr, spec.r2_insert, bar)
    r1 = None
    if spec.paired:
        h1 = hdr.copy()
        h1[:, p + 16] = ord("1")
        r1 = record_block(h1, spec.r1_len, None)
    return r1, r2


REFORMAT_BARCODE = "ACGTTGCAAC"


def reformat_block(spec: LaneSpec, block: int, n_pairs: int):
    """The lane's block as ONE sample's reads the way a splitting-barcode run leaves them: the header ends with
    " n:N:0:BARCODE" instead of "/n" (a tenth of the reads with one base of it changed).  Input of `mgikit reformat`."""
    r1, r2 = make_block(spec, block, n_pairs)
    rng = np.random.default_rng([spec.seed, block, 11])
    hl = 1 + len(FLOWCELL) + 2 + 4 + 4 + 8 + 2 + 1
    tail = np.frombuffer((":N:0:" + REFORMAT_BARCODE + "\n").encode(), dtype=np.uint8)
    bad = np.nonzero(rng.random(n_pairs) < 0.1)[0]
    at = rng.integers(0, len(REFORMAT_BARCODE), size=bad.size)

    def widen(flat):
        if flat is None:
            return None
        rec = flat.reshape(n_pairs, -1)
        out = np.empty((n_pairs, rec.shape[1] + tail.size - 1), dtype=np.uint8)
        out[:, :hl - 3] = rec[:, :hl - 3]
        out[:, hl - 3] = ord(" ")
        out[:, hl - 2] = rec[:, hl - 2]          # the read number
        out[:, hl - 1:hl - 1 + tail.size] = tail
        out[bad, hl - 1 + 5 + at] = ord("N")
        out[:, hl - 1 + tail.size:] = rec[:, hl:]
        return out.reshape(-1)

    return widen(r1), widen(r2)


def reformat_run_spec(spec: LaneSpec, umi_length: int = 0, **kw) -> RunSpec:
    return RunSpec(samples=[Sample("S001", REFORMAT_BARCODE, ".")], template="i7" + str(len(REFORMAT_BARCODE)), paired=spec.paired,
                   mismatches=5, l_position=L_POSITION, illumina_prefix=f"@INS:20240101:{FLOWCELL}:", reformat=True,
                   reformat_umi_length=umi_length, reformat_barcode=REFORMAT_BARCODE, **kw)


def record_bytes(spec: LaneSpec) -> Tuple[int, int]:
    hl = 1 + len(FLOWCELL) + 2 + 4 + 4 + 8 + 2 + 1
    BL = sum(n for _, n in _template_items(spec.template))
    r2 = hl + 2 * (spec.r2_insert + BL + 1) + 2
    r1 = hl + 2 * (spec.r1_len + 1) + 2 if spec.paired else 0
    return r1, r2
