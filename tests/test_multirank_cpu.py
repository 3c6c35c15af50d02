"""N > 1 path on CPU (world_size 2, gloo): chunks are sharded round-robin over ranks
(chunk k -> rank k mod N), every rank keeps private counter tables, ONE
all-reduce merges them, and appending every bin's segments in chunk order restores the
single-process output byte for byte.  The per-rank engine here is the oracle (there is no
GPU in this container); the GPU path does the same through mgk_nccl_allreduce_counters."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mgikit_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_CHUNKS, N_PAIRS, WORLD = 6, 700, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_chunks(hp, ks):
    out = {}
    for k in ks:
        r1, r2 = synth.make_block(synth.CONFIG2, k, N_PAIRS)
        res = hp.process(r2, r1, seq=k)
        out[k] = ([res.segment(2, b) for b in range(hp.total)], [res.segment(1, b) for b in range(hp.total)])
    return out


def _worker(rank, port, q):
    import ctypes
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import make_hotpath
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "liboracle.so"))
    spec = synth.run_spec(synth.CONFIG2, max_chunk_bytes=4 << 20, max_chunk_records=N_PAIRS + 64)
    hp = make_hotpath(lib, "orc_", spec)
    mine = _run_chunks(hp, range(rank, N_CHUNKS, WORLD))
    stats, mism = hp.counters()
    t = torch.from_numpy(np.concatenate([stats.ravel(), mism.ravel()]).astype(np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)          # == mgk_nccl_allreduce_counters on the GPU path
    gathered = [None] * WORLD
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put((t.numpy(), gathered))
    dist.barrier()
    dist.destroy_process_group()
    hp.close()


def test_two_ranks_equal_one(oracle_lib):
    from util import make_hotpath
    spec = synth.run_spec(synth.CONFIG2, max_chunk_bytes=4 << 20, max_chunk_records=N_PAIRS + 64)
    hp = make_hotpath(oracle_lib, "orc_", spec)
    serial = _run_chunks(hp, range(N_CHUNKS))
    s_stats, s_mism = hp.counters()
    hp.close()

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    merged, gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(merged, np.concatenate([s_stats.ravel(), s_mism.ravel()]).astype(np.int64))
    by_chunk = {}
    for part in gathered:
        by_chunk.update(part)
    assert sorted(by_chunk) == list(range(N_CHUNKS))
    total = len(serial[0][0])
    for which in (0, 1):
        for b in range(total):   # appending in chunk order == what one process writes
            assert b"".join(by_chunk[k][which][b] for k in range(N_CHUNKS)) == b"".join(serial[k][which][b] for k in range(N_CHUNKS))
