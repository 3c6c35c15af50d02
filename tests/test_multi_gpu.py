"""The N > 1 path (SURVEY §8 a13 / e).

CPU (-m "not gpu"): the host pump's round-robin + out-of-place merge, with the oracle as
the per-device engine.  GPU (-m gpu, needs >= 2 devices): the product CLI with --gpus 2
against the reference's goldens, and two contexts on two devices merged by
mgk_allreduce_counters (ncclCommInitAll + ncclAllReduce) against the oracle's tables."""
import numpy as np
import pytest

import multigpu_check
from mgikit_b200 import synth
from util import make_hotpath


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def test_oracle_cli_two_devices_equals_goldens(oracle_cli):
    """Host side only: chunk k -> context k mod 2, segments appended in chunk order, merged counter tables."""
    done = multigpu_check.cli_matches_goldens(oracle_cli, 2)
    assert set(done) == {"cfg5", "large-pe"}


def test_oracle_merge_is_out_of_place_and_repeatable(oracle_lib):
    """ReportManager::update semantics of the merge: sum of the shares, shares untouched, can be repeated."""
    lane = synth.CONFIG2
    spec = synth.run_spec(lane, max_chunk_bytes=4 << 20, max_chunk_records=2100)
    a, b = make_hotpath(oracle_lib, "orc_", spec), make_hotpath(oracle_lib, "orc_", spec)
    for k in range(4):
        r1, r2 = synth.make_block(lane, k, 2000)
        (a if k % 2 == 0 else b).process(r2, r1, seq=k)
    sa, ma = a.counters()
    sb, mb = b.counters()
    import ctypes
    arr = (ctypes.c_void_p * 2)(a.ctx, b.ctx)
    for _ in range(2):
        assert oracle_lib.orc_allreduce_counters(arr, 2) == 0
        for h in (a, b):
            s, m = h.reduced_counters()
            np.testing.assert_array_equal(s, sa + sb)
            np.testing.assert_array_equal(m, ma + mb)
    np.testing.assert_array_equal(a.counters()[0], sa)
    np.testing.assert_array_equal(b.counters()[1], mb)
    a.close()
    b.close()


@pytest.mark.gpu
@pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_cuda_cli_two_gpus_equals_goldens(product_cli):
    done = multigpu_check.cli_matches_goldens(product_cli, 2)
    assert set(done) == {"cfg5", "large-pe"}


@pytest.mark.gpu
@pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_contexts_allreduce_equals_sum_and_oracle(oracle_lib, gpu_lib):
    """mgk_allreduce_counters over two devices: reduced tables == element-wise sum of the two shares == the oracle's
    tables over all chunks; the shares themselves stay (out of place), a second merge gives the same."""
    import ctypes
    import dataclasses
    lane = synth.CONFIG5
    n = 3000
    base = synth.run_spec(lane, max_chunk_bytes=4 << 20, max_chunk_records=n + 64)
    g = [make_hotpath(gpu_lib, "mgk_", dataclasses.replace(base, device=d)) for d in (0, 1)]
    o = make_hotpath(oracle_lib, "orc_", base)
    for k in range(6):
        r1, r2 = synth.make_block(lane, k, n)
        g[k % 2].process(r2, r1, seq=k)
        o.process(r2, r1, seq=k)
    shares = [h.counters() for h in g]
    arr = (ctypes.c_void_p * 2)(g[0].ctx, g[1].ctx)
    so, mo = o.counters()
    for _ in range(2):
        assert gpu_lib.mgk_allreduce_counters(arr, 2) == 0, gpu_lib.mgk_last_error(g[0].ctx)
        for h in g:
            s, m = h.reduced_counters()
            np.testing.assert_array_equal(s, shares[0][0] + shares[1][0])
            np.testing.assert_array_equal(m, shares[0][1] + shares[1][1])
            np.testing.assert_array_equal(s, so)
            np.testing.assert_array_equal(m, mo)
    for h, sh in zip(g, shares):
        np.testing.assert_array_equal(h.counters()[0], sh[0])
    for h in g:
        h.close()
    o.close()
