"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle
on the same seeded inputs — every result field and both counter tables bit for bit."""
import dataclasses

import numpy as np
import pytest

import cases
from mgikit_b200 import synth
from mgikit_b200.capi import MGK_IN_DEVICE, MGK_OUT_DEVICE
from util import assert_same_counters, assert_same_result, device_copy, make_hotpath, run_both

pytestmark = pytest.mark.gpu
ALL = cases.all_cases()


@pytest.mark.parametrize("name,spec,chunks", ALL, ids=[c[0] for c in ALL])
def test_cuda_matches_oracle(oracle_lib, gpu_lib, name, spec, chunks):
    run_both(oracle_lib, "orc_", gpu_lib, "mgk_", spec, chunks, what=name)


def test_newline_scan_stage(oracle_lib, gpu_lib):
    """a1: K1 against memchr on adversarial layouts (runs of newlines, none, tile boundaries)."""
    spec = synth.run_spec(synth.CONFIG2, max_chunk_bytes=8 << 20, max_chunk_records=1 << 16)
    g = make_hotpath(gpu_lib, "mgk_", spec)
    o = make_hotpath(oracle_lib, "orc_", spec)
    rng = np.random.default_rng(3)
    bufs = [np.zeros(0, np.uint8), np.full(1, 10, np.uint8), np.full(5000, 10, np.uint8), np.full(40000, 65, np.uint8)]
    for n in (1, 15, 16, 17, 63, 64, 65, 2047, 2048, 2049, 4095, 4096, 4097, 16383, 16384, 16385, 16384 * 3 + 5, 1 << 20):
        b = rng.integers(32, 127, size=n, dtype=np.uint8)
        b[rng.random(n) < 0.02] = 10
        bufs.append(b)
    # every byte value (the flag arithmetic must not take 0x0b, 0x8a, 0x00 ... for a newline), dense newlines
    b = rng.integers(0, 256, size=300000, dtype=np.uint8)
    b[rng.random(b.size) < 0.05] = 10
    bufs.append(b)
    bufs.append(np.tile(np.array([10, 11, 0x8A, 9, 0x0A, 0x1A, 0x2A, 0x80, 0, 0x7F, 0xFF, 10, 10], np.uint8), 4001))
    # one 4 KiB stage per scan warp of the whole grid, one byte either side (every range holds one tile or none)
    k = g.scan_ranges() * 4096 if hasattr(g, "scan_ranges") else 1776 * 4096
    for n in (k - 1, k, k + 1, k + 4096 + 3):
        b = rng.integers(32, 127, size=n, dtype=np.uint8)
        b[rng.random(n) < 0.011] = 10
        bufs.append(b)
    r1, r2 = synth.make_block(synth.CONFIG2, 0, 5000)
    bufs += [r1, r2]
    for b in bufs:
        np.testing.assert_array_equal(g.scan_newlines(b, capacity=b.size + 16), o.scan_newlines(b, capacity=b.size + 16))
    g.close()
    o.close()


@pytest.mark.parametrize("lane", [synth.CONFIG2, synth.CONFIG3, synth.CONFIG4], ids=lambda l: l.name)
def test_match_stage_random_barcodes(oracle_lib, gpu_lib, lane):
    """a6: find_matching_sample on 200k barcodes: exact, near, N, lowercase, random."""
    spec = synth.run_spec(lane)
    g = make_hotpath(gpu_lib, "mgk_", spec)
    o = make_hotpath(oracle_lib, "orc_", spec)
    _, r2 = synth.make_block(dataclasses.replace(lane, frac_random=0.2, frac_n=0.05, frac_one_mm=0.2, frac_two_mm=0.1), 3, 200000)
    _, rb = synth.record_bytes(lane)
    BL = sum(int(it[2:]) for it in lane.template.split(":"))
    recs = r2.reshape(-1, rb)
    seq_end = 32 + lane.r2_insert + BL
    bars = recs[:, seq_end - BL:seq_end].copy()
    bars[::97, 0] = ord("a")
    sg, mg = g.match_barcodes(bars)
    so, mo = o.match_barcodes(bars)
    np.testing.assert_array_equal(sg, so)
    np.testing.assert_array_equal(mg, mo)
    g.close()
    o.close()


def test_device_resident_and_double_buffered(oracle_lib, gpu_lib):
    """Chunks already in HBM (MGK_IN_DEVICE), two in flight, outputs left on the device and
    fetched with torch: same bytes as the host path and as the oracle."""
    import torch
    lane = synth.CONFIG2
    n = 4000
    spec = synth.run_spec(lane, max_chunk_bytes=4 << 20, max_chunk_records=n + 64, n_slots=2)
    g = make_hotpath(gpu_lib, "mgk_", spec)
    o = make_hotpath(oracle_lib, "orc_", spec)
    blocks = [synth.make_block(lane, b, n) for b in range(4)]
    dev = [(device_copy(r2), device_copy(r1)) for r1, r2 in blocks]
    results = {}
    for k in range(len(blocks)):
        (_, a2), (_, a1) = dev[k]
        g.submit(k, a2, a1, MGK_IN_DEVICE | MGK_OUT_DEVICE)
        if k >= 1:
            results[k - 1] = g.collect(device_out=True)
            _fetch(results[k - 1])
            g.release(k - 1)
    results[len(blocks) - 1] = g.collect(device_out=True)
    _fetch(results[len(blocks) - 1])
    g.release(len(blocks) - 1)
    for k, (r1, r2) in enumerate(blocks):
        assert_same_result(o.process(r2, r1, seq=k), results[k], f"device chunk {k}")
    assert_same_counters(o, g, "device-resident")
    g.close()
    o.close()


def _fetch(res):
    import ctypes

    import torch
    for which in ("r2", "r1"):
        n = getattr(res, f"out_{which}_bytes")
        ptr = getattr(res, f"out_{which}_ptr")
        host = torch.empty(n, dtype=torch.uint8)
        if n:
            torch.cuda.synchronize()
            cudart = ctypes.CDLL("libcudart.so.12")
            rc = cudart.cudaMemcpy(ctypes.c_void_p(host.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(n), 2)
            assert rc == 0
        setattr(res, f"out_{which}", host.numpy().tobytes())


def test_full_size_properties(gpu_lib):
    """Size-independent properties at a BASELINE-sized chunk (1M pairs of config 2): every input
    record lands in exactly one bin, per-bin order is input order, counters add up."""
    lane = synth.CONFIG2
    n = 1 << 20
    r1, r2 = synth.make_block(lane, 0, n)
    spec = synth.run_spec(lane, max_chunk_bytes=512 << 20, max_chunk_records=n + 64, n_slots=1)
    g = make_hotpath(gpu_lib, "mgk_", spec)
    res = g.process(r2, r1)
    assert res.n_records == n and res.consumed_r2 == r2.size and res.consumed_r1 == r1.size
    stats, mism = g.counters()
    assert int(mism[:, 1:].sum()) == n
    out1 = np.frombuffer(res.out_r1, np.uint8)
    assert int((out1 == 10).sum()) == 4 * n                       # every record kept its 4 lines
    # read numbers inside every bin are strictly increasing (order preserved)
    total = spec_total = len(spec.samples) + 2
    und = spec_total - 2
    seg = res.segment(1, und)
    heads = [l for l in seg.split(b"\n")[0::4] if l]
    nums = [int(h[21:29]) for h in heads]
    assert nums == sorted(nums) and len(set(nums)) == len(nums)
    seg0 = res.segment(1, 0)
    ids = [int(l.split(b":")[4]) for l in seg0.split(b"\n")[0::4] if l]
    assert ids == sorted(ids)
    assert int(res.seg_len_r1.sum()) == res.out_r1_bytes and int(res.seg_len_r2.sum()) == res.out_r2_bytes
    g.close()


@pytest.mark.parametrize("seed", range(24))
def test_cuda_matches_oracle_random_cases(oracle_lib, gpu_lib, seed):
    """The randomised cases of test_fuzz_cpu.py through the kernels."""
    from test_fuzz_cpu import _case
    spec, chunks = _case(1000 + seed)
    run_both(oracle_lib, "orc_", gpu_lib, "mgk_", spec, chunks, what=f"fuzz seed {seed}: {spec.template} m={spec.mismatches} mode={spec.out_mode}")


@pytest.mark.parametrize("seed", range(16))
def test_cuda_matches_oracle_random_mixed_libraries(oracle_lib, gpu_lib, seed):
    from test_fuzz_cpu import _mixed_case
    spec, chunks = _mixed_case(5000 + seed)
    run_both(oracle_lib, "orc_", gpu_lib, "mgk_", spec, chunks,
             what=f"mixed fuzz seed {seed}: m={spec.mismatches} comp={spec.comprehensive_scan} mode={spec.out_mode}")


def test_device_input_without_slack_is_refused(gpu_lib):
    """MGK_IN_DEVICE contract (include/mgikit_gpu.h): the allocation must stay readable MGK_DEVICE_INPUT_SLACK bytes
    past the text; a buffer that ends with its text is refused, not over-read."""
    import ctypes

    from mgikit_b200.capi import MgkError
    lane = synth.CONFIG2
    r1, r2 = synth.make_block(lane, 0, 512)
    spec = synth.run_spec(lane, max_chunk_bytes=1 << 20, max_chunk_records=600)
    g = make_hotpath(gpu_lib, "mgk_", spec)
    cudart = ctypes.CDLL("libcudart.so.12")
    size = (r2.size + 511) & ~511                     # cudaMalloc'd ranges are what cuMemGetAddressRange reports
    p = ctypes.c_void_p()
    assert cudart.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(size)) == 0
    try:
        (_, a1) = device_copy(r1)
        tight = (p.value + size - r2.size) & ~15      # text ends within 16 bytes of the allocation's end
        with pytest.raises(MgkError) as e:
            g.submit(0, (tight, p.value + size - tight), a1, MGK_IN_DEVICE)
        assert e.value.code == -1 and "readable bytes" in str(e.value)
    finally:
        cudart.cudaFree(p)
        g.close()
