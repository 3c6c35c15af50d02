"""Seeded chunk-level cases shared by the CPU (oracle vs emulation) and GPU
(oracle vs CUDA) parity tests."""
from __future__ import annotations

import dataclasses

import numpy as np

from mgikit_b200 import synth
from mgikit_b200.plan import RunSpec, Sample


def _lane_chunks(lane: synth.LaneSpec, n_pairs: int, n_chunks: int = 1):
    out = []
    for b in range(n_chunks):
        r1, r2 = synth.make_block(lane, b, n_pairs)
        out.append((r2, r1))
    return out


def config_cases(n_pairs=3000):
    """BASELINE.json configs 2-5 at a size the oracle finishes in seconds."""
    cases = []
    for lane in (synth.CONFIG2, synth.CONFIG3, synth.CONFIG4, synth.CONFIG5):
        spec = synth.run_spec(lane, max_chunk_bytes=8 << 20, max_chunk_records=n_pairs + 64)
        cases.append((lane.name, spec, _lane_chunks(lane, n_pairs, 2)))
    return cases


def flag_cases(n_pairs=1200):
    lane = dataclasses.replace(synth.CONFIG2, n_samples=24, frac_random=0.08, frac_n=0.02)
    lane_umi = dataclasses.replace(synth.CONFIG3, n_samples=24, frac_random=0.08, frac_n=0.02)
    chunks = _lane_chunks(lane, n_pairs)
    chunks_umi = _lane_chunks(lane_umi, n_pairs)
    kw = dict(max_chunk_bytes=4 << 20, max_chunk_records=n_pairs + 64)
    base = synth.run_spec(lane, **kw)
    base_umi = synth.run_spec(lane_umi, **kw)
    R = dataclasses.replace
    cases = [
        ("mgi-format", R(base, out_mode=1), chunks),
        ("mgi-full-header", R(base, out_mode=2), chunks),
        ("mgi-full-header-umi", R(base_umi, out_mode=2), chunks_umi),
        ("illumina-umi-m0", R(base_umi, mismatches=0), chunks_umi),
        ("keep-barcode", R(base, keep_barcode=True), chunks),
        ("report-level-0", R(base, report_level=0), chunks),
        ("report-level-1", R(base, report_level=1), chunks),
        ("per-index-error-m2", R(base, mismatches=2, per_index_error=True), chunks),
        ("m0", R(base, mismatches=0), chunks),
        ("m3", R(base, mismatches=3), chunks),
        ("validate-clean", R(base, validate=True), chunks),
        ("single-end-from-r2", R(base, paired=False), [(c[0], None) for c in chunks]),
    ]
    # duplicated sample ids share an output bin (src/sample_manager.rs:692-730)
    dup = R(base, samples=[Sample(f"D{k % 5}", s.i7, s.i5) for k, s in enumerate(base.samples)])
    cases.append(("duplicate-sample-ids", dup, chunks))
    return cases


def _fastq(records):
    return np.frombuffer(b"".join(records), dtype=np.uint8)


def edge_cases():
    """Hand-made chunks for the corners the reference's code has."""
    samples = [Sample("A", "AAAAAAAA", "CCCCCCCC"), Sample("B", "AAAAAAAT", "GGGGGGGG"),
               Sample("C", "ACGTACGT", "TGCATGCA"), Sample("D", "ACGTACGA", "TGCATGCA")]
    kw = dict(template="i78:i58", l_position=5, illumina_prefix="@I:R:FC01:", max_chunk_bytes=1 << 20,
              max_chunk_records=4096)

    def pair(n, bar, hdr=None, q="I", insert="ACGTACGTAC", r1seq="TTTTGGGGCCCCAAAA"):
        h = hdr if hdr is not None else f"@FC01L1C{n % 1000:03d}R{(n * 7) % 1000:03d}{n:07d}"
        s2 = insert + bar
        r2 = f"{h}/2\n{s2}\n+\n{q * len(s2)}\n".encode()
        r1 = f"{h}/1\n{r1seq}\n+\n{'5' * len(r1seq)}\n".encode()
        return r2, r1

    def chunk(pairs):
        return _fastq([p[0] for p in pairs]), _fastq([p[1] for p in pairs])

    cases = []
    bars = ["AAAAAAAACCCCCCCC", "AAAAAAATGGGGGGGG", "AAAAAAAAGGGGGGGG", "ACGTACGTTGCATGCA", "ACGTACGATGCATGCA",
            "ACGTACGCTGCATGCA", "NAAAAAAACCCCCCCC", "aAAAAAAACCCCCCCC", "AAAAAAAACCCCCCCN", "NNNNNNNNNNNNNNNN",
            "ACGTACGTTGCATGCa", "TTTTTTTTTTTTTTTT", "ACGTACGNTGCATGCA"]
    pairs = [pair(i + 1, b) for i, b in enumerate(bars * 3)]
    for m in (0, 1, 2):
        cases.append((f"nearest-set-rule-m{m}", RunSpec(samples=samples, mismatches=m, **kw), [chunk(pairs)]))
    cases.append(("nearest-set-rule-m1-per-index", RunSpec(samples=samples, mismatches=1, per_index_error=True, **kw),
                  [chunk(pairs)]))
    # header number fields: leading zeros, all zeros, long flowcell-like read numbers
    hdrs = ["@FC01L1C001R00100000000", "@FC01L1C000R00000000100", "@FC01L1C120R10500012345", "@FC01L1C012R03400000014",
            "@FC01L1C001R001", "@FC01L1C100R0909", "@FC01L1C010R0010000000000000000000000000007"]
    hp = [pair(i, "AAAAAAAACCCCCCCC", hdr=h) for i, h in enumerate(hdrs)]
    cases.append(("header-zero-stripping", RunSpec(samples=samples, **kw), [chunk(hp)]))
    # rewritten headers longer than a staging slot take the byte-store path; mixed with normal ones
    long_hdrs = [f"@FC01L1C{i % 7:03d}R{i % 5:03d}" + "0" * (i % 4) + "1234567890" * (3 + i % 9) for i in range(40)]
    lp = [pair(i, bars[i % len(bars)], hdr=h) for i, h in enumerate(long_hdrs)]
    cases.append(("header-longer-than-slot", RunSpec(samples=samples, **kw), [chunk(lp + pairs[:9] + lp[:7])]))
    cases.append(("header-longer-than-slot-full", RunSpec(samples=samples, out_mode=2, **kw), [chunk(lp + pairs[:9])]))
    # ragged records: every read its own lengths
    rng = np.random.default_rng(5)
    rag = []
    for i in range(400):
        ins = "".join("ACGT"[v] for v in rng.integers(0, 4, size=int(rng.integers(0, 90))))
        r1s = "".join("ACGT"[v] for v in rng.integers(0, 4, size=int(rng.integers(1, 200))))
        rag.append(pair(i + 1, bars[i % len(bars)], insert=ins, r1seq=r1s, q="?"))
    cases.append(("ragged-lengths", RunSpec(samples=samples, **kw), [chunk(rag)]))
    cases.append(("ragged-lengths-mgi", RunSpec(samples=samples, out_mode=1, **kw), [chunk(rag)]))
    # long reads: bodies of many 16-byte units (the movers' loop past the units a lane takes in registers), the seam
    # between sequence and qualities anywhere in them
    lng = []
    for i in range(300):
        ins = "".join("ACGT"[v] for v in rng.integers(0, 4, size=int(rng.integers(180, 900))))
        r1s = "".join("ACGT"[v] for v in rng.integers(0, 4, size=int(rng.integers(180, 1100))))
        lng.append(pair(i + 1, bars[i % len(bars)], insert=ins, r1seq=r1s, q="F"))
    cases.append(("long-reads", RunSpec(samples=samples, **kw), [chunk(lng)]))
    cases.append(("long-reads-keep-barcode-mgi", RunSpec(samples=samples, out_mode=1, keep_barcode=True, **kw), [chunk(lng)]))
    # very short reads: bodies shorter than one unit, seams inside the head or tail bytes
    tiny = []
    for i in range(300):
        ins = "".join("ACGT"[v] for v in rng.integers(0, 4, size=int(rng.integers(0, 6))))
        r1s = "".join("ACGT"[v] for v in rng.integers(0, 4, size=int(rng.integers(1, 9))))
        tiny.append(pair(i + 1, bars[i % len(bars)], insert=ins, r1seq=r1s, q="#", hdr=f"@F{i % 10}L1C{i % 1000:03d}R{(i * 3) % 1000:03d}{i}"))
    cases.append(("tiny-reads", RunSpec(samples=samples, **kw), [chunk(tiny)]))
    cases.append(("tiny-reads-mgi", RunSpec(samples=samples, out_mode=1, **kw), [chunk(tiny)]))
    # empty chunk, one record, trailing partial record in either stream
    one = chunk(pairs[:1])
    cases.append(("single-record", RunSpec(samples=samples, **kw), [one]))
    cases.append(("empty-chunk", RunSpec(samples=samples, **kw), [(np.zeros(0, np.uint8), np.zeros(0, np.uint8))]))
    r2, r1 = chunk(pairs[:5])
    cases.append(("partial-tail-r2", RunSpec(samples=samples, **kw), [(r2[:-7], r1)]))
    cases.append(("partial-tail-r1", RunSpec(samples=samples, **kw), [(r2, r1[:-3])]))
    cases.append(("no-newline-at-all", RunSpec(samples=samples, **kw), [(r2[:20], r1[:20])]))
    # barcode-only barcode read: samples get no barcode-read file (src/lib.rs:1724-1727)
    bo = [pair(i + 1, b, insert="") for i, b in enumerate(bars)]
    cases.append(("barcode-only-read2", RunSpec(samples=samples, read2_has_sequence=False, **kw), [chunk(bo)]))
    # --validate failures: first offending record and its kind
    bad_q = pairs[:6] + [pair(7, bars[0], q="~")] + pairs[7:9]
    cases.append(("validate-bad-quality", RunSpec(samples=samples, validate=True, **kw), [chunk(bad_q)]))
    bad_b = pairs[:3] + [pair(4, bars[0], insert="ACGTXCGTAC")] + pairs[4:6]
    cases.append(("validate-bad-base", RunSpec(samples=samples, validate=True, **kw), [chunk(bad_b)]))
    r2x, r1x = chunk(pairs[:6])
    r1y = r1x.copy()
    r1y[np.nonzero(r1y == ord("C"))[0][3]] = ord("D")   # header of a later record differs between the reads
    cases.append(("validate-header-mismatch", RunSpec(samples=samples, validate=True, **kw), [(r2x, r1y)]))
    r2z = r2x.copy()
    r2z[0] = ord(">")
    cases.append(("validate-header-at", RunSpec(samples=samples, validate=True, **kw), [(r2z, r1x)]))
    return cases


def template_cases(n_pairs=800):
    """Mixed libraries: several templates, comprehensive scan on and off, single index + UMI."""
    rng = np.random.default_rng(11)
    s = [Sample("M1", "TGCTGTGA", "CACGATTC", "i78:um8:i58", "1", "1"),
         Sample("M2", "CCACATTG", "TACTCCAG", "i78:um8:i58", "1", "1"),
         Sample("M3", "TAGTGCCA", "CGTCAAGA", "i78:um8:i58", "0", "0"),
         Sample("M4", "TCCTGTGA", "CACGACCC", "i58:i78:um8", "0", "1"),
         Sample("M5", "AGAACCAG", ".", "um16:i78", "1", "."),
         Sample("M6", "GGAACCAG", ".", "um16:i78", "0", ".")]
    pool = ["TCACAGCA", "CAATGTGG", "TAGTGCCA", "TCCTGTGA", "CTGGTTCT", "GGAACCAG", "GAATCGTG", "CTGGAGTA", "CGTCAAGA",
            "GGGTCGTG", "AGAACCAG", "CACGATTC"]
    recs2, recs1 = [], []
    for i in range(n_pairs):
        parts = [pool[int(rng.integers(0, len(pool)))] for _ in range(3)]
        if rng.random() < 0.3:
            k = int(rng.integers(0, 3))
            p = list(parts[k])
            p[int(rng.integers(0, 8))] = "ACGTN"[int(rng.integers(0, 5))]
            parts[k] = "".join(p)
        bar = "".join(parts)
        h = f"@FC02L1C{i % 900:03d}R{i % 77:03d}{i:08d}"
        recs2.append(f"{h}/2\nACGTACGTACGT{bar}\n+\n{'F' * (12 + 24)}\n".encode())
        recs1.append(f"{h}/1\nGGGGCCCCAAAATTTT\n+\n{'9' * 16}\n".encode())
    chunk = (_fastq(recs2), _fastq(recs1))
    kw = dict(l_position=5, illumina_prefix="@I:R:FC02:", max_chunk_bytes=1 << 20, max_chunk_records=4096)
    cases = []
    for m in (0, 1, 2):
        for comp in (False, True):
            cases.append((f"mixed-library-m{m}-comp{int(comp)}",
                          RunSpec(samples=s, mismatches=m, comprehensive_scan=comp, **kw), [chunk]))
    cases.append(("mixed-library-full-header", RunSpec(samples=s, out_mode=2, comprehensive_scan=True, **kw), [chunk]))
    return cases


def all_cases():
    return config_cases() + flag_cases() + edge_cases() + template_cases()
