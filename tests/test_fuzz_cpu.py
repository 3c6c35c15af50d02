"""Randomised chunk-level cases (CPU): the device logic compiled by g++ against the oracle on run descriptions and
records drawn at random -- template shapes, index lengths, flags, header digits, read lengths, barcode damage.  The
fixed cases of cases.py pin what the reference's tests pin; this looks for what nobody thought of."""
import dataclasses

import numpy as np
import pytest

from mgikit_b200.plan import RunSpec, Sample
from util import run_both

TEMPLATES = [("i7{a}:i5{b}", True, False), ("i5{b}:i7{a}", True, False), ("i7{a}:um{u}:i5{b}", True, True),
             ("um{u}:i7{a}", False, True), ("i7{a}", False, False), ("i7{a}:--3:i5{b}", True, False)]


def _rand_seq(rng, n, alphabet="ACGT"):
    return "".join(alphabet[v] for v in rng.integers(0, len(alphabet), size=n))


def _case(seed):
    rng = np.random.default_rng(seed)
    shape, dual, has_umi = TEMPLATES[int(rng.integers(0, len(TEMPLATES)))]
    a, b, u = int(rng.integers(6, 13)), int(rng.integers(6, 13)), int(rng.integers(4, 10))
    tmpl = shape.format(a=a, b=b, u=u)
    n_samples = int(rng.integers(2, 40))
    seen, samples = set(), []
    while len(samples) < n_samples:
        i7, i5 = _rand_seq(rng, a), (_rand_seq(rng, b) if dual else ".")
        if (i7, i5) in seen or (not dual and any(s.i7 == i7 for s in samples)):
            continue
        seen.add((i7, i5))
        sid = f"S{len(samples) % max(1, n_samples - int(rng.integers(0, 3)))}"   # now and then two rows share an id
        samples.append(Sample(sid, i7, i5))
    paired = bool(rng.random() < 0.8)
    out_mode = int(rng.integers(0, 3))
    l_pos = int(rng.integers(3, 12))
    spec = RunSpec(samples=samples, template=tmpl, i7_rc=bool(rng.random() < 0.3), i5_rc=bool(dual and rng.random() < 0.3),
                   paired=paired, out_mode=out_mode, keep_barcode=bool(rng.random() < 0.2),
                   report_level=int(rng.integers(0, 3)), mismatches=int(rng.integers(0, 4)),
                   per_index_error=bool(rng.random() < 0.3), l_position=l_pos,
                   illumina_prefix="@" + _rand_seq(rng, int(rng.integers(1, 30)), "ABCxyz019") + ":",
                   max_chunk_bytes=1 << 20, max_chunk_records=4096)
    # what the read carries: the sheet's index, reverse-complemented where the flag says the sheet is
    comp = str.maketrans("ACGT", "TGCA")
    rc = lambda s: s.translate(comp)[::-1]
    items = [(it[:2], int(it[2:])) for it in tmpl.split(":")]
    flow = _rand_seq(rng, l_pos - 1, "ABCDEFGHV0123456789")
    recs2, recs1 = [], []
    for i in range(int(rng.integers(1, 400))):
        s = samples[int(rng.integers(0, n_samples))]
        parts = []
        for kind, n in items:
            if kind == "i7":
                parts.append(rc(s.i7) if spec.i7_rc else s.i7)
            elif kind == "i5":
                parts.append(rc(s.i5) if spec.i5_rc else s.i5)
            else:
                parts.append(_rand_seq(rng, n))
        bar = list("".join(parts))
        r = rng.random()
        if r < 0.25:      # substitutions
            for _ in range(int(rng.integers(1, 4))):
                bar[int(rng.integers(0, len(bar)))] = "ACGTN"[int(rng.integers(0, 5))]
        elif r < 0.30:    # something the dictionary alphabet does not have
            bar[int(rng.integers(0, len(bar)))] = "acgtnX."[int(rng.integers(0, 7))]
        elif r < 0.35:
            bar = list(_rand_seq(rng, len(bar)))
        bar = "".join(bar)
        ins = _rand_seq(rng, int(rng.integers(0, 260)))
        num = str(int(rng.integers(0, 10 ** int(rng.integers(1, 12))))).zfill(int(rng.integers(1, 14)))
        h = f"@{flow}L{int(rng.integers(1, 5))}C{int(rng.integers(0, 1000)):03d}R{int(rng.integers(0, 1000)):03d}{num}"
        q2 = "".join(chr(int(v)) for v in rng.integers(33, 74, size=len(ins) + len(bar)))
        recs2.append(f"{h}/2\n{ins}{bar}\n+\n{q2}\n".encode())
        r1s = _rand_seq(rng, int(rng.integers(1, 300)))
        q1 = "".join(chr(int(v)) for v in rng.integers(33, 74, size=len(r1s)))
        recs1.append(f"{h}/1\n{r1s}\n+\n{q1}\n".encode())
    r2 = np.frombuffer(b"".join(recs2), np.uint8)
    r1 = np.frombuffer(b"".join(recs1), np.uint8) if paired else None
    return spec, [(r2, r1)]


@pytest.mark.parametrize("seed", range(60))
def test_random_case_device_logic_matches_oracle(oracle_lib, emu_lib, seed):
    spec, chunks = _case(1000 + seed)
    run_both(oracle_lib, "orc_", emu_lib, "emu_", spec, chunks, what=f"fuzz seed {seed}: {spec.template} m={spec.mismatches} mode={spec.out_mode}")


def _mixed_case(seed):
    """Several templates of one barcode length in one sheet (per-sample template and rc columns), with and without
    --comprehensive-scan; the reads mix every template's layout and damaged barcodes."""
    rng = np.random.default_rng(seed)
    BL = 24
    shapes = ["i78:um8:i58", "i58:i78:um8", "um16:i78", "i712:i512", "um8:i78:i58", "i78:--8:i58"]
    picks = [shapes[int(v)] for v in rng.choice(len(shapes), size=int(rng.integers(2, 4)), replace=False)]
    pool = [_rand_seq(rng, 12) for _ in range(10)]   # few index strings, shared between templates: collisions on purpose
    samples, seen = [], set()
    for k in range(int(rng.integers(3, 16))):
        t = picks[int(rng.integers(0, len(picks)))]
        items = [(it[:2], int(it[2:])) for it in t.split(":")]
        l7 = next(n for kd, n in items if kd == "i7")
        l5 = next((n for kd, n in items if kd == "i5"), 0)
        i7 = pool[int(rng.integers(0, len(pool)))][:l7]
        i5 = pool[int(rng.integers(0, len(pool)))][:l5] if l5 else "."
        if (t, i7, i5) in seen or (not l5 and any(x[0] == t and x[1] == i7 for x in seen)):
            continue
        seen.add((t, i7, i5))
        samples.append(Sample(f"M{k}", i7, i5, t, str(int(rng.integers(0, 2))), str(int(rng.integers(0, 2))) if l5 else "."))
    if not samples:
        samples.append(Sample("M0", pool[0][:8], pool[1][:8], "i78:um8:i58", "0", "0"))
    spec = RunSpec(samples=samples, paired=True, out_mode=int(rng.integers(0, 3)), comprehensive_scan=bool(rng.random() < 0.5),
                   mismatches=int(rng.integers(0, 3)), per_index_error=bool(rng.random() < 0.3), l_position=5,
                   report_level=2, illumina_prefix="@I:R:FC02:", max_chunk_bytes=1 << 20, max_chunk_records=4096)
    recs2, recs1 = [], []
    for i in range(int(rng.integers(50, 500))):
        bar = list("".join(pool[int(rng.integers(0, len(pool)))][:8] for _ in range(3)))
        if rng.random() < 0.4:
            for _ in range(int(rng.integers(1, 3))):
                bar[int(rng.integers(0, BL))] = "ACGTN"[int(rng.integers(0, 5))]
        bar = "".join(bar)
        h = f"@FC02L1C{i % 900:03d}R{i % 77:03d}{i:08d}"
        ins = _rand_seq(rng, int(rng.integers(0, 40)))
        recs2.append(f"{h}/2\n{ins}{bar}\n+\n{'F' * (len(ins) + BL)}\n".encode())
        recs1.append(f"{h}/1\nGGGGCCCCAAAATTTT\n+\n{'9' * 16}\n".encode())
    return spec, [(np.frombuffer(b"".join(recs2), np.uint8), np.frombuffer(b"".join(recs1), np.uint8))]


@pytest.mark.parametrize("seed", range(40))
def test_random_mixed_library_device_logic_matches_oracle(oracle_lib, emu_lib, seed):
    spec, chunks = _mixed_case(5000 + seed)
    run_both(oracle_lib, "orc_", emu_lib, "emu_", spec, chunks,
             what=f"mixed fuzz seed {seed}: m={spec.mismatches} comp={spec.comprehensive_scan} mode={spec.out_mode}")


def _reformat_case(seed):
    """`reformat` mode: headers of every shape (blank-separated tails, raw "/n", blanks in odd places, several blanks,
    long tails), UMI lengths from none to the whole read, with and without a sample barcode, both output modes."""
    rng = np.random.default_rng(seed)
    paired = bool(rng.random() < 0.8)
    l_pos = int(rng.integers(3, 12))
    bcl = int(rng.integers(0, 3)) * int(rng.integers(4, 13))          # 0 (none) in a third of the cases
    barcode = _rand_seq(rng, bcl) + ("+" + _rand_seq(rng, bcl) if bcl and rng.random() < 0.3 else "")
    umi = int(rng.integers(0, 3)) * int(rng.integers(1, 12))
    spec = RunSpec(samples=[Sample("only", "ACGTACGT", ".")], template="i78", paired=paired,
                   out_mode=0 if rng.random() < 0.85 else 1, report_level=int(rng.integers(0, 3)), mismatches=5, l_position=l_pos,
                   validate=bool(rng.random() < 0.15), illumina_prefix="@" + _rand_seq(rng, int(rng.integers(1, 60)), "ABCxyz019") + ":",
                   max_chunk_bytes=1 << 20, max_chunk_records=4096, reformat=True, reformat_umi_length=umi, reformat_barcode=barcode)
    flow = _rand_seq(rng, l_pos - 1, "ABCDEFGHV0123456789")
    strict = bool(rng.random() < 0.7)   # only headers both implementations must accept
    recs2, recs1 = [], []
    for i in range(int(rng.integers(1, 300))):
        num = str(int(rng.integers(0, 10 ** int(rng.integers(1, 12))))).zfill(int(rng.integers(1, 14)))
        h = f"@{flow}L{int(rng.integers(1, 5))}C{int(rng.integers(0, 1000)):03d}R{int(rng.integers(0, 1000)):03d}{num}"
        hb = list(barcode)
        if hb and rng.random() < 0.4:
            for _ in range(int(rng.integers(1, 7))):
                hb[int(rng.integers(0, len(hb)))] = "ACGTN+"[int(rng.integers(0, 6))]
        hb = "".join(hb) if barcode else _rand_seq(rng, int(rng.integers(0, 20)))
        shape = int(rng.integers(0, 8)) if not strict else int(rng.integers(0, 3))
        def tail(n):
            if shape in (0, 1):
                return f" {n}:N:0:{hb}"
            if shape == 2:
                return f"/{n}"
            if shape == 3:
                return f" {n}"                                      # blank three bytes before the end
            if shape == 4:
                return f" {n}:N:0:{hb} " + _rand_seq(rng, int(rng.integers(0, 230)), "xyz: ")   # may break the barcode length
            if shape == 5:
                return f"{n} "                                      # blank in the last two bytes
            if shape == 6:
                return f" {n}:N:0:" + _rand_seq(rng, int(rng.integers(0, 30)))                  # barcode of another length
            return ""                                               # no read number at all
        t2, t1 = tail(2), tail(1)
        if shape == 4:
            t1 = t2.replace(" 2:", " 1:", 1)
        n2 = int(rng.integers(umi, umi + 200))
        s2 = _rand_seq(rng, n2, "ACGTN")
        q2 = "".join(chr(int(v)) for v in rng.integers(33, 74, size=n2))
        recs2.append(f"{h}{t2}\n{s2}\n+\n{q2}\n".encode())
        n1 = int(rng.integers(1, 300))
        recs1.append(f"{h}{t1}\n{_rand_seq(rng, n1)}\n+\n{''.join(chr(int(v)) for v in rng.integers(33, 74, size=n1))}\n".encode())
    r2 = np.frombuffer(b"".join(recs2), np.uint8)
    r1 = np.frombuffer(b"".join(recs1), np.uint8) if paired else None
    return spec, [(r2, r1)]


@pytest.mark.parametrize("seed", range(80))
def test_random_reformat_case_device_logic_matches_oracle(oracle_lib, emu_lib, seed):
    """Both accept and agree byte for byte, or both refuse the chunk (headers the reference itself would panic on)."""
    from mgikit_b200.capi import MgkError
    from util import assert_same_counters, assert_same_result, make_hotpath
    spec, chunks = _reformat_case(5000 + seed)
    ha, hb = make_hotpath(oracle_lib, "orc_", spec), make_hotpath(emu_lib, "emu_", spec)
    try:
        r2, r1 = chunks[0]
        res, err = [], []
        for hp in (ha, hb):
            try:
                res.append(hp.process(r2, r1, seq=0))
                err.append(None)
            except MgkError as e:
                res.append(None)
                err.append(e.code)
        assert (err[0] is None) == (err[1] is None), f"seed {seed}: oracle {err[0]} vs device logic {err[1]}"
        if err[0] is None:
            assert_same_result(res[0], res[1], f"reformat fuzz seed {seed}")
            if not res[0].validate_error:
                assert_same_counters(ha, hb, f"reformat fuzz seed {seed}")
    finally:
        ha.close()
        hb.close()
