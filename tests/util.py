"""Shared helpers of the parity tests: run two implementations of the C ABI on
the same chunk and compare every field of the result bit for bit."""
from __future__ import annotations

import numpy as np

from mgikit_b200.capi import HotPath
from mgikit_b200.plan import RunSpec, build_config


def make_hotpath(lib, prefix: str, spec: RunSpec) -> HotPath:
    cfg, keep = build_config(spec)
    return HotPath(lib, prefix, cfg, keepalive=keep)


def assert_same_result(a, b, what=""):
    assert (a.validate_error, a.validate_record if a.validate_error else 0) == \
           (b.validate_error, b.validate_record if b.validate_error else 0), f"{what}: validate verdict differs"
    if a.validate_error:
        return  # the reference panics at that record; nothing after the verdict is defined
    assert a.n_records == b.n_records, f"{what}: n_records {a.n_records} vs {b.n_records}"
    assert (a.consumed_r2, a.consumed_r1) == (b.consumed_r2, b.consumed_r1), f"{what}: consumed bytes differ"
    np.testing.assert_array_equal(a.seg_len_r2, b.seg_len_r2, err_msg=f"{what}: barcode-read segment lengths")
    np.testing.assert_array_equal(a.seg_len_r1, b.seg_len_r1, err_msg=f"{what}: paired-read segment lengths")
    np.testing.assert_array_equal(a.seg_off_r2, b.seg_off_r2, err_msg=f"{what}: barcode-read segment offsets")
    np.testing.assert_array_equal(a.seg_off_r1, b.seg_off_r1, err_msg=f"{what}: paired-read segment offsets")
    assert a.out_r2_bytes == b.out_r2_bytes and a.out_r1_bytes == b.out_r1_bytes
    if a.out_r2 != b.out_r2:
        x = np.frombuffer(a.out_r2, np.uint8)
        y = np.frombuffer(b.out_r2, np.uint8)
        i = int(np.nonzero(x != y)[0][0])
        raise AssertionError(f"{what}: barcode-read output differs at byte {i}: {a.out_r2[max(0, i - 60):i + 40]!r} vs {b.out_r2[max(0, i - 60):i + 40]!r}")
    if (a.out_r1 or b"") != (b.out_r1 or b""):
        x = np.frombuffer(a.out_r1, np.uint8)
        y = np.frombuffer(b.out_r1, np.uint8)
        i = int(np.nonzero(x != y)[0][0])
        raise AssertionError(f"{what}: paired-read output differs at byte {i}: {a.out_r1[max(0, i - 60):i + 40]!r} vs {b.out_r1[max(0, i - 60):i + 40]!r}")
    np.testing.assert_array_equal(np.sort(a.unassigned), np.sort(b.unassigned), err_msg=f"{what}: unassigned list")


def assert_same_counters(hp_a: HotPath, hp_b: HotPath, what=""):
    sa, ma = hp_a.counters()
    sb, mb = hp_b.counters()
    np.testing.assert_array_equal(ma, mb, err_msg=f"{what}: sample_mismatches table")
    np.testing.assert_array_equal(sa, sb, err_msg=f"{what}: sample_statistics table")


def run_both(lib_a, pre_a, lib_b, pre_b, spec: RunSpec, chunks, what=""):
    """chunks: list of (r2, r1) byte arrays.  Compares every chunk and the final counters."""
    ha = make_hotpath(lib_a, pre_a, spec)
    hb = make_hotpath(lib_b, pre_b, spec)
    try:
        for k, (r2, r1) in enumerate(chunks):
            ra = ha.process(r2, r1, seq=k)
            rb = hb.process(r2, r1, seq=k)
            assert_same_result(ra, rb, f"{what} chunk {k}")
            if ra.validate_error:
                return
        assert_same_counters(ha, hb, what)
    finally:
        ha.close()
        hb.close()


def device_copy(arr):
    """A uint8 array copied to the current CUDA device with the readable slack MGK_IN_DEVICE asks for
    (include/mgikit_gpu.h: MGK_DEVICE_INPUT_SLACK bytes past the text); returns (tensor that owns it, (ptr, len))."""
    import torch
    t = torch.empty(arr.size + 64, dtype=torch.uint8, device="cuda")
    t[:arr.size].copy_(torch.from_numpy(arr))
    return t, (t.data_ptr(), arr.size)
