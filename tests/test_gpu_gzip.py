"""MGK_OUT_GZIP (SURVEY §8 f-2 on the device): every bin's segment comes back as a run of gzip members whose
decompressed bytes are exactly the text segment the oracle produces -- the comparison the reference's own tests
make for .gz files (tests/integration_test.rs:44-53).  Also the framing: BGZF-compatible members of <= 32 KiB text."""
import dataclasses
import gzip
import struct
import zlib

import numpy as np
import pytest

import cases
from mgikit_b200 import synth
from mgikit_b200.capi import MGK_OUT_GZIP
from util import assert_same_counters, make_hotpath

pytestmark = pytest.mark.gpu


def _members(blob: bytes):
    """Splits a run of BGZF members using their BSIZE fields; yields (member bytes, text)."""
    at = 0
    while at < len(blob):
        assert blob[at:at + 4] == b"\x1f\x8b\x08\x04", "every member is gzip/deflate with an extra field"
        xlen = struct.unpack_from("<H", blob, at + 10)[0]
        assert xlen == 6 and blob[at + 12:at + 16] == b"BC\x02\x00"
        size = struct.unpack_from("<H", blob, at + 16)[0] + 1
        member = blob[at:at + size]
        d = zlib.decompressobj(31)
        text = d.decompress(member)
        assert d.eof and not d.unused_data, "BSIZE spans exactly one member"
        yield member, text
        at += size


def _check(oracle_lib, gpu_lib, spec, chunks, what):
    g = make_hotpath(gpu_lib, "mgk_", spec)
    o = make_hotpath(oracle_lib, "orc_", spec)
    try:
        for k, (r2, r1) in enumerate(chunks):
            want = o.process(r2, r1, seq=k)
            got = g.process(r2, r1, seq=k, flags=MGK_OUT_GZIP)
            assert got.validate_error == want.validate_error
            if want.validate_error:   # the reference panics at that record; nothing after the verdict is defined
                return
            assert got.n_records == want.n_records
            for which in (2, 1):
                for b in range(g.total):
                    z, text = got.segment(which, b), want.segment(which, b)
                    assert (len(z) == 0) == (len(text) == 0), f"{what} chunk {k} bin {b}: empty segments stay empty"
                    if not text:
                        continue
                    assert gzip.decompress(z) == text, f"{what} chunk {k} read {which} bin {b}: gunzip(members) != text segment"
                    parts = list(_members(z))
                    assert b"".join(t for _, t in parts) == text
                    assert all(len(t) <= 32768 for _, t in parts) and all(len(t) == 32768 for _, t in parts[:-1])
            np.testing.assert_array_equal(np.sort(got.unassigned), np.sort(want.unassigned))
        assert_same_counters(o, g, what)
    finally:
        g.close()
        o.close()


ALL = cases.all_cases()
PICK = [c for i, c in enumerate(ALL) if i % 4 == 0]


@pytest.mark.parametrize("name,spec,chunks", PICK, ids=[c[0] for c in PICK])
def test_gzip_segments_decompress_to_the_oracle_text(oracle_lib, gpu_lib, name, spec, chunks):
    _check(oracle_lib, gpu_lib, spec, chunks, name)


@pytest.mark.parametrize("lane,n", [(synth.CONFIG2, 30000), (synth.CONFIG3, 12000), (synth.CONFIG4, 50000), (synth.CONFIG5, 9000)],
                         ids=["cfg2", "cfg3", "cfg4", "cfg5"])
def test_gzip_baseline_lanes(oracle_lib, gpu_lib, lane, n):
    """Segments that span many 32 KiB blocks (few samples take most reads) and the many tiny ones of 384 samples."""
    skew = dataclasses.replace(lane, frac_random=0.3)
    chunks = []
    for k in range(2):
        r1, r2 = synth.make_block(skew, k, n)
        chunks.append((r2, r1))
    spec = synth.run_spec(lane, max_chunk_bytes=64 << 20, max_chunk_records=n + 64, n_slots=2)
    _check(oracle_lib, gpu_lib, spec, chunks, lane.name)


def test_gzip_of_incompressible_text_is_stored(oracle_lib, gpu_lib):
    """Quality bytes over the whole range and random headers do not shrink under an order-0 code: the block is stored,
    never expanded past its slot."""
    lane = dataclasses.replace(synth.CONFIG2, n_samples=4)
    n = 3000
    r1, r2 = synth.make_block(lane, 0, n)
    rng = np.random.default_rng(5)
    _, rb = synth.record_bytes(lane)
    recs = r2.reshape(n, rb).copy()
    q0 = 32 + 170 + 3
    recs[:, q0:q0 + 150] = rng.integers(33, 74, size=(n, 150), dtype=np.uint8)
    recs[:, 32:32 + 150] = rng.choice(np.frombuffer(b"ACGTNacgtn", np.uint8), size=(n, 150))
    spec = synth.run_spec(lane, max_chunk_bytes=8 << 20, max_chunk_records=n + 64)
    _check(oracle_lib, gpu_lib, spec, [(recs.reshape(-1), r1)], "wide alphabet")
