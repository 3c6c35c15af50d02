"""The input side of the pump, on the CPU: the product's own inflate
(mgikit_b200/csrc/host/inflate.cpp) and member-parallel gunzip (gz_reader.cpp) against zlib --
every block type, window size and strategy, streams suspended at arbitrary output boundaries,
truncated and bit-flipped streams, multi-member files with the gzip magic planted inside the
data, padding and garbage after the last member.  Decoded bytes must equal zlib's, errors must
be errors (flate2's MultiGzDecoder, /root/reference/src/file_utils.rs:231-233, behaves like zlib)."""
import ctypes
import gzip
import os
import zlib

import numpy as np
import pytest

from mgikit_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hc():
    p = os.path.join(ROOT, "oracle", "_ref", "libhostcodec.so")
    if not os.path.exists(p):
        import subprocess
        subprocess.run(["make", "-C", ROOT, "oracle"], check=True, capture_output=True)
    lib = ctypes.CDLL(p)
    lib.hc_crc32.restype = ctypes.c_uint32
    lib.hc_crc32.argtypes = [ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t]
    lib.hc_inflate_raw.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                   ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
    lib.hc_gunzip_file2.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p,
                                    ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t),
                                    ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]
    lib.hc_gunzip_file.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                   ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t),
                                   ctypes.c_char_p, ctypes.c_size_t]
    return lib


def _inflate(lib, z, cap, piece=0):
    out = ctypes.create_string_buffer(cap + 400)
    p, c = ctypes.c_size_t(), ctypes.c_size_t()
    st = lib.hc_inflate_raw(z, len(z), out, cap + 300, piece, ctypes.byref(p), ctypes.byref(c))
    return st, out.raw[:p.value], c.value


def _raw_deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
    return c.compress(data) + c.flush()


def _fastq(n, block=0):
    return synth.make_block(synth.CONFIG2, block, n)[1].tobytes()


def test_crc32_matches_zlib(hc):
    rng = np.random.default_rng(1)
    for n in [0, 1, 7, 15, 16, 63, 64, 65, 100, 1000, 4096, 65537, 1 << 20]:
        b = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert hc.hc_crc32(0, b, n) == zlib.crc32(b)
        assert hc.hc_crc32(12345, b, n) == zlib.crc32(b, 12345)


def test_crc32_combine_matches_zlib(hc):
    hc.hc_crc32_combine.restype = ctypes.c_uint32
    hc.hc_crc32_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
    rng = np.random.default_rng(7)
    for la, lb in [(0, 0), (5, 0), (0, 7), (100, 1), (1, 100000), (123457, 765432), (1 << 20, (1 << 21) + 3)]:
        a = rng.integers(0, 256, la, dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, lb, dtype=np.uint8).tobytes()
        assert hc.hc_crc32_combine(zlib.crc32(a), zlib.crc32(b), lb) == zlib.crc32(a + b), (la, lb)


def test_inflate_equals_zlib_on_every_kind_of_stream(hc):
    rng = np.random.default_rng(2)
    datas = [b"", b"a", b"abc" * 1000, bytes(100000), _fastq(4000), rng.integers(0, 256, size=150000, dtype=np.uint8).tobytes(),
             rng.integers(65, 69, size=200000, dtype=np.uint8).tobytes(), (b"ACGT" * 7 + b"\n") * 9000]
    for d in datas:
        for level in (0, 1, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                for wb in (-15, -9):
                    z = _raw_deflate(d, level, strat, wb)
                    for piece in (0, 70000, 1000):   # one buffer; suspended and resumed at buffer ends
                        st, o, c = _inflate(hc, z + b"TRAILINGJUNK1234567", len(d), piece)
                        assert st == 0 and o == d and c == len(z), (len(d), level, strat, wb, piece, st, len(o), c, len(z))


def test_inflate_flags_truncated_and_damaged_streams(hc):
    rng = np.random.default_rng(3)
    d = _fastq(600)
    z = _raw_deflate(d, 6)
    for cut in list(range(0, 150)) + [len(z) // 2, len(z) - 1]:
        st, _, _ = _inflate(hc, z[:cut], len(d))
        assert st in (2, 3), (cut, st)
    for k in range(400):
        zz = bytearray(z)
        i = int(rng.integers(0, len(zz)))
        zz[i] ^= 1 << int(rng.integers(0, 8))
        st, o, _ = _inflate(hc, bytes(zz), 2 * len(d))
        try:
            ref = zlib.decompressobj(-15).decompress(bytes(zz))
            ok = True
        except zlib.error:
            ok = False
        if ok:     # zlib got through (possibly stopping early without error): whatever we finished must agree
            assert st in (0, 1, 3) and (st != 0 or o == ref), (k, st)
        else:
            assert st != 0, (k, st)


def _gunzip_file(lib, path, threads, cap, max_buffered=64 << 20):
    out = ctypes.create_string_buffer(cap + 16)
    p, nl, m = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    err = ctypes.create_string_buffer(512)
    rc = lib.hc_gunzip_file(path.encode(), threads, max_buffered, out, cap, ctypes.byref(p), ctypes.byref(nl), ctypes.byref(m), err, 512)
    return rc, out.raw[:p.value], nl.value, m.value, err.value.decode()


def _member(data, level=1, name=b""):
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = c.compress(data) + c.flush()
    flg = 8 if name else 0
    hdr = b"\x1f\x8b\x08" + bytes([flg]) + b"\0\0\0\0\0\x03" + (name + b"\0" if name else b"")
    return hdr + body + zlib.crc32(data).to_bytes(4, "little") + (len(data) & 0xffffffff).to_bytes(4, "little")


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_gunzip_multi_member_files_in_parallel(hc, tmp_path, threads):
    rng = np.random.default_rng(4)
    # members of very different sizes; some hold the gzip magic + plausible flags INSIDE their (stored) data: candidates
    # the reader must never take for members
    decoy = b"\x1f\x8b\x08\x00" + bytes(14)
    parts = [_fastq(3000, b) for b in range(5)] + [b"", decoy * 500, rng.integers(0, 256, size=300000, dtype=np.uint8).tobytes() + decoy,
                                                  _fastq(40000, 9), b"x"]
    blob = b"".join(_member(p, level=(0 if decoy in p else 1), name=(b"lane.fq" if i % 3 == 0 else b"")) for i, p in enumerate(parts))
    want = b"".join(parts)
    path = str(tmp_path / "multi.fq.gz")
    with open(path, "wb") as f:
        f.write(blob)
    assert gzip.decompress(blob) == want
    for max_buffered in (64 << 20, 1):      # the smallest buffer budget forces the workers to wait for the consumer
        rc, got, nl, members, err = _gunzip_file(hc, path, threads, len(want), max_buffered)
        assert rc == 0, err
        assert got == want and nl == want.count(b"\n") and members == len(parts)


def test_gunzip_single_member_streams_through_small_buffers(hc, tmp_path):
    want = _fastq(60000)          # 22 MB: several 4 MiB pieces with the window carried between them
    path = str(tmp_path / "one.fq.gz")
    with gzip.open(path, "wb", compresslevel=1) as f:
        f.write(want)
    rc, got, nl, members, err = _gunzip_file(hc, path, 4, len(want), 1)
    assert rc == 0 and got == want and members == 1 and nl == want.count(b"\n"), err


def test_gunzip_refuses_what_zlib_refuses(hc, tmp_path):
    good = _member(_fastq(2000)) + _member(_fastq(2000, 1))
    cases = {"empty": (b"", 0), "zero padded": (good + bytes(512), 0), "not gzip": (b"@read\nACGT\n+\nIIII\n", 1),
             "truncated": (good[:-9], 1), "cut inside the second member": (good[:len(good) * 3 // 4], 1),
             "garbage after": (good + b"garbage that is no member", 1),
             "bad crc": (good[:-8] + bytes([good[-8] ^ 1]) + good[-7:], 1),
             "bad length": (good[:-1] + bytes([good[-1] ^ 1]), 1)}
    for name, (blob, want_rc) in cases.items():
        path = str(tmp_path / "case.gz")
        with open(path, "wb") as f:
            f.write(blob)
        rc, got, _, _, err = _gunzip_file(hc, path, 3, 4 << 20)
        assert rc == want_rc, (name, rc, err)
        if want_rc == 0 and blob:
            assert got == gzip.decompress(good)


def test_gunzip_reference_fixtures(hc):
    """Every .fq.gz the reference ships as test input decodes to what zlib makes of it."""
    base = os.path.join(ROOT, "tests", "golden", "testing_data", "input")
    n = 0
    for d, _, files in os.walk(base):
        for f in files:
            if f.endswith(".gz"):
                with open(os.path.join(d, f), "rb") as fh:
                    want = gzip.decompress(fh.read())
                rc, got, nl, _, err = _gunzip_file(hc, os.path.join(d, f), 2, len(want) + 16)
                assert rc == 0 and got == want and nl == want.count(b"\n"), (f, err)
                n += 1
    assert n >= 10


def _gunzip_parts(lib, path, threads, cap, parallel_min, part_bytes):
    out = ctypes.create_string_buffer(cap + 16)
    p, nl, m = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    err = ctypes.create_string_buffer(512)
    rc = lib.hc_gunzip_file2(path.encode(), threads, 256 << 20, parallel_min, part_bytes, out, cap, ctypes.byref(p), ctypes.byref(nl),
                             ctypes.byref(m), err, 512)
    return rc, out.raw[:p.value], nl.value, m.value, err.value.decode()


@pytest.mark.parametrize("level,strategy", [(1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                            (1, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY)],
                         ids=["l1", "l6", "l9", "huffman-only", "fixed-blocks", "stored"])
def test_one_member_decoded_in_parts_equals_zlib(hc, tmp_path, level, strategy):
    """A long member is taken up at several places at once (block finder + markers for the unknown 32 KiB in front of a
    part, resolved when the part before it is through): same bytes as one pass.  Fixed and stored blocks give the
    finder nothing to hold on to: the part before simply runs on."""
    want = _fastq(40000) + bytes(range(256)) * 300 + _fastq(20000, 3)
    c = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
    blob = c.compress(want) + c.flush()
    path = str(tmp_path / "long.fq.gz")
    with open(path, "wb") as f:
        f.write(blob)
    for threads, part in ((2, 1 << 20), (5, 200000), (8, 65536)):
        rc, got, nl, members, err = _gunzip_parts(hc, path, threads, len(want), 1 << 18, part)
        assert rc == 0, err
        assert got == want and nl == want.count(b"\n") and members == 1, (threads, part)


def test_parts_inside_a_multi_member_file_and_damage(hc, tmp_path):
    big = _fastq(50000, 1)
    small = [_fastq(1500, b) for b in range(3)]
    blob = _member(small[0]) + _member(big, level=6) + _member(small[1]) + _member(small[2])
    want = small[0] + big + small[1] + small[2]
    path = str(tmp_path / "mix.fq.gz")
    with open(path, "wb") as f:
        f.write(blob)
    rc, got, _, members, err = _gunzip_parts(hc, path, 6, len(want), 1 << 18, 300000)
    assert rc == 0 and got == want and members == 4, err
    # a flipped bit deep inside the long member: an error, never silently different bytes
    rng = np.random.default_rng(11)
    for k in range(12):
        bad = bytearray(blob)
        i = len(_member(small[0])) + int(rng.integers(2000, len(_member(big, level=6)) - 2000))
        bad[i] ^= 1 << int(rng.integers(0, 8))
        with open(path, "wb") as f:
            f.write(bytes(bad))
        rc, got, _, _, err = _gunzip_parts(hc, path, 6, len(want) + (1 << 20), 1 << 18, 300000)
        try:
            ref = gzip.decompress(bytes(bad))
        except Exception:
            ref = None
        assert (rc == 0) == (ref is not None), (k, rc, err)
        if ref is not None:
            assert got == ref
    with open(path, "wb") as f:
        f.write(blob[:len(blob) // 2])
    rc, _, _, _, err = _gunzip_parts(hc, path, 6, len(want), 1 << 18, 300000)
    assert rc == 1 and "end of file" in err


def test_members_the_scanner_does_not_list_are_still_taken(hc, tmp_path):
    """The scanner lists likely starts only (XFL in {0, 2, 4}); a member with another XFL byte is reached by the chain and
    accepted on its header alone."""
    parts = [_fastq(2500, b) for b in range(4)]
    blob = b""
    for i, p in enumerate(parts):
        m = bytearray(_member(p))
        m[8] = 7 if i % 2 else 0
        blob += bytes(m)
    path = str(tmp_path / "xfl.fq.gz")
    with open(path, "wb") as f:
        f.write(blob)
    rc, got, _, members, err = _gunzip_file(hc, path, 3, sum(map(len, parts)))
    assert rc == 0 and got == b"".join(parts) and members == 4, err
