#!/usr/bin/env python
"""Benchmark of the demultiplex hot path (BASELINE.json: read pairs/s demultiplexed,
bit-exact; HBM GB/s as a fraction of peak).

A step = one pass of the whole hot path (record scan -> match -> bin scan -> scatter+QC)
over one batch of synthetic read pairs: `--chunks-per-step` chunks of `--pairs` pairs.
N = 1 runs BASELINE config 2 (96 samples, dual 10 bp indexes, 150 bp pairs, m=1), N > 1
config 5 (384 samples; chunks sharded by rank, one NCCL all-reduce of the counter tables).
`value` is measured with the batch already resident in HBM (CUDA events on the launching
streams); `e2e` goes through the public C ABI with pinned HOST buffers, H2D and D2H copies
inside the timed region, next to a copy-only leg over the same buffers (`copy_ceiling`);
`e2e_files` is the drop-in CLI on .fq.gz files next to the reference binary on the same files.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...                     (the reference's CPU binary)
"""
import argparse
import json
import os
import shutil
import signal
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "read_pairs_per_second_demultiplexed"
UNIT = "read pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 21, help="read pairs per chunk (1.5 GB of text: > 10x L2)")
    ap.add_argument("--chunks-per-step", type=int, default=32,
                    help="chunks per step and GPU: long enough steps that the device-timed span is >= 0.5 s")
    ap.add_argument("--e2e-chunk-pairs", type=int, default=1 << 20, help="end to end, a chunk goes through the C ABI in pieces of at most this many pairs")
    ap.add_argument("--cpu-sample-pairs", type=int, default=400000)
    ap.add_argument("--file-pairs", type=int, default=10_000_000, help="read pairs of the .fq.gz lane of the file-level leg (0: skip)")
    ap.add_argument("--file-timeout", type=int, default=420, help="seconds the file-level leg may take before it is given up")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--in-flight", type=int, default=2, help="chunks in flight in the resident leg (= slots of the context)")
    return ap.parse_args()


def workload(lane, gpus):
    from mgikit_b200 import synth
    which = {"cfg2": "configs[1]", "cfg5": "configs[4]"}[lane.name[:4]]
    r1_rec, r2_rec = synth.record_bytes(lane)
    return (f"BASELINE {which}: synthetic MGI paired-end lane, {lane.n_samples} samples, dual 10bp indexes ({lane.template}), "
            f"150bp+150bp(+20bp barcode), mismatches={lane.mismatches}, Illumina-format output"
            + ("" if gpus == 1 else f", sharded over {gpus} GPUs with an NCCL all-reduce of the report tables")), r1_rec, r2_rec


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for k, nme in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU reference arm
def reference_binary():
    for p in (os.path.join(ROOT, "oracle", "_ref", "mgikit"), "/root/reference/bins/mgikit"):
        if os.path.exists(p) and os.access(p, os.X_OK):
            try:
                if subprocess.run([p, "--version"], capture_output=True, timeout=20).returncode == 0:
                    return p, "reference"
            except Exception:
                pass
    p = os.path.join(ROOT, "oracle", "_ref", "mgikit_oracle")
    if not os.path.exists(p):
        subprocess.run(["make", "-C", ROOT, "oracle"], check=True, capture_output=True)
    return p, "port"


def cpu_run(lane, n_pairs, steps, warmup):
    """Times `mgikit demultiplex` (the reference's own binary when it runs here, else the
    oracle port) on a bounded gz sample of the workload, with every host core."""
    import lanes
    binary, kind = reference_binary()
    cores = os.cpu_count() or 1
    d = tempfile.mkdtemp(prefix="mgk_cpu_")
    try:
        r1, r2, sheet = lanes.write_lane(lane, n_pairs, d)
        times = []
        for it in range(warmup + steps):
            out = os.path.join(d, "out")
            args = lanes.demux_args(lane, r1, r2, sheet, out)
            args[args.index("-t") + 1] = str(cores)
            t0 = time.perf_counter()
            lanes.run_binary(binary, args)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
            shutil.rmtree(out, ignore_errors=True)
        return {"value": n_pairs * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": kind,
                "sample": f"{n_pairs} read pairs of the same synthetic lane as .fq.gz (gzip -1), `mgikit demultiplex -t {cores} "
                          f"--writing-buffer-size 4194304`, gunzip+demultiplex+gzip, {len(times)} timed runs",
                "seconds_per_run": sum(times) / len(times)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ------------------------------------------------------------------ resident runs (any lane)
def resident_run(hp, dev_args, n_chunks, flags, depth=2):
    """n_chunks submissions of the resident block, `depth` in flight; returns (per-kernel ms rows, last result)."""
    rows, res, done = [], None, 0
    for k in range(n_chunks):
        hp.submit(k, dev_args[0], dev_args[1], flags)
        if k - done + 1 >= depth:
            res = hp.collect(device_out=True)
            rows.append(hp.last_timings())
            hp.release(done)
            done += 1
    while done < n_chunks:
        res = hp.collect(device_out=True)
        rows.append(hp.last_timings())
        hp.release(done)
        done += 1
    return rows, res


def serial_run(hp, dev_args, n_chunks, flags):
    """One chunk in flight at a time: per-kernel CUDA-event times that no other launch overlaps."""
    rows = []
    for k in range(n_chunks):
        hp.submit(k, dev_args[0], dev_args[1], flags)
        hp.collect(device_out=True)
        rows.append(hp.last_timings())
        hp.release(k)
    return rows


T_START = time.perf_counter()


def progress(what):
    """One line per finished leg on stderr (rank 0 only says where a run is, should it ever stop)."""
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench] {time.perf_counter() - T_START:7.1f} s  {what}", file=sys.stderr, flush=True)


def short_lane_line(lane, n, device, peak, reformat=False):
    """other_configs: a short resident run of another BASELINE configuration (same CUDA-event method); reformat: the
    lane as one sample's reads through the one-sample mode of the path (`mgikit reformat`, 8-base UMI)."""
    import numpy as np
    import torch
    from mgikit_b200 import synth
    from mgikit_b200.capi import MGK_IN_DEVICE, MGK_OUT_DEVICE, load_gpu_library
    from util import device_copy, make_hotpath
    r1, r2 = synth.reformat_block(lane, 0, n) if reformat else synth.make_block(lane, 0, n)
    in_bytes = r2.size + (r1.size if r1 is not None else 0)
    caps = dict(max_chunk_bytes=max(r1.size if r1 is not None else 0, r2.size) + 4096, max_chunk_records=n + 64, n_slots=2, device=device)
    spec = synth.reformat_run_spec(lane, umi_length=8, **caps) if reformat else synth.run_spec(lane, **caps)
    hp = make_hotpath(load_gpu_library(), "mgk_", spec)
    try:
        d2, a2 = device_copy(r2)
        d1, a1 = device_copy(r1) if r1 is not None and r1.size else (None, None)
        dev_args = (a2, a1)
        flags = MGK_IN_DEVICE | MGK_OUT_DEVICE
        resident_run(hp, dev_args, 4, flags)
        k_ms = np.array(serial_run(hp, dev_args, 6, flags)[1:])
        hp.span_ms()
        n_chunks = 24
        _, res = resident_run(hp, dev_args, n_chunks, flags)
        span = hp.span_ms()
        out_bytes = res.out_r2_bytes + res.out_r1_bytes
        emit_ms = float(k_ms[:, 3].mean())
        achieved = (in_bytes + out_bytes) / (emit_ms * 1e-3) / 1e9
        return {"lane": lane.name + ("-as-one-sample-reformat-umi8" if reformat else ""), "units_per_chunk": n, "unit": "read pairs/s" if lane.paired else "reads/s",
                "value": n * n_chunks / (span * 1e-3), "bytes_per_unit": {"in": in_bytes / n, "out": out_bytes / n},
                "kernel_ms": {"scan_newlines": float(k_ms[:, 0].mean()), "match_plan": float(k_ms[:, 1].mean()),
                              "bin_scan": float(k_ms[:, 2].mean()), "emit_qc": emit_ms, "chunk": float(k_ms[:, 4].mean())},
                "roofline": {"kernel": "k_emit", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak},
                "whole_path_frac": ((in_bytes + out_bytes) * n_chunks / (span * 1e-3) / 1e9) / peak}
    finally:
        hp.close()


def main():
    a = parse()
    from mgikit_b200 import synth
    lane = synth.CONFIG2 if a.gpus == 1 else synth.CONFIG5
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    what, r1_rec, r2_rec = workload(lane, a.gpus)
    cps = max(1, a.chunks_per_step)
    step_pairs = a.pairs * cps
    config = {"workload": what, "pairs_per_step_per_gpu": step_pairs, "chunks_per_step": cps, "pairs_per_chunk": a.pairs,
              "input_bytes_per_pair": r1_rec + r2_rec,
              "l2_policy": f"every chunk streams {(r1_rec + r2_rec) * a.pairs >> 20} MiB in + about as much out per GPU (>> 126 MB L2); "
                           "no explicit flush needed",
              "parallelism": f"{a.gpus} GPU(s), read chunks sharded by rank, one NCCL allreduce of the counter tables"}
    if a.gpus > 1:   # N = 1 runs config 2 (the headline), N > 1 config 5 (BASELINE's multi-GPU configuration)
        config["weak_scaling_reference"] = "single_gpu_same_lane of this line (rank 0 alone, same box), or the N=1 line's other_configs.cfg5"

    if a.impl == "reference":
        if rank != 0:
            return
        res = cpu_run(lane, a.cpu_sample_pairs, max(1, a.steps), min(a.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": res["seconds_per_run"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from mgikit_b200.capi import CLI_BIN, MGK_IN_DEVICE, MGK_OUT_DEVICE, MGK_OUT_GZIP, load_gpu_library
    from util import device_copy, make_hotpath

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    try:   # run (and first-touch the pinned buffers) on the cores next to this rank's GPU: on a two-socket box the
           # end-to-end leg is bound by the host side of the copies
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        print(f"[bench] rank {rank}: {len(os.sched_getaffinity(0))} of {len(all_cpus)} cores (those next to GPU {local})", file=sys.stderr)
    except Exception as e:   # not fatal: the default placement is what every run before this had
        print(f"[bench] rank {rank}: default core placement ({type(e).__name__}: {e})", file=sys.stderr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = a.pairs
    r1, r2 = synth.make_block(lane, rank, n)          # every rank its own block of the lane (weak scaling)
    in_bytes = r1.size + r2.size
    spec = synth.run_spec(lane, max_chunk_bytes=max(r1.size, r2.size) + 4096, max_chunk_records=n + 64, n_slots=max(3, a.in_flight), device=local)
    lib = load_gpu_library()
    hp = make_hotpath(lib, "mgk_", spec)
    if world > 1:   # NCCL communicator of the library itself, id shared through torch.distributed
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(hp.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        hp.nccl_join(bytes(uid.cpu().numpy().tobytes()), rank, world)
        hp.nccl_allreduce_counters()   # connection setup happens on the first collective: not part of any timing below

    (d2, a2), (d1, a1) = device_copy(r2), device_copy(r1)
    dev_args = (a2, a1)
    RES = MGK_IN_DEVICE | MGK_OUT_DEVICE

    def merged_tables_check(expected_reads):
        """The one collective of the path: every rank's share of the two report tables, summed by the library's NCCL
        all-reduce, must equal the element-wise sum of the shares as gathered independently with torch.distributed
        (ReportManager::update, src/report_manager.rs:377-435), on every rank."""
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hp.nccl_allreduce_counters()
        ms = (time.perf_counter() - t0) * 1e3
        stats, mism = hp.counters()
        rs, rm = hp.reduced_counters()
        mine = torch.from_numpy(np.concatenate([stats.ravel(), mism.ravel()]).astype(np.int64)).cuda()
        shares = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(shares, mine)
        want = torch.stack(shares).sum(dim=0).cpu().numpy().astype(np.uint64)
        got = np.concatenate([rs.ravel(), rm.ravel()])
        assert np.array_equal(got, want), f"rank {rank}: all-reduced report tables differ from the sum of the per-rank tables"
        assert int(rm[:, 1:].sum()) == expected_reads, "merged counter tables do not add up to the reads processed"
        assert int(mism[:, 1:].sum()) == expected_reads // world, "this rank's counter tables do not add up to its reads"
        return ms

    # ---- resident (kernel path) timing
    hp.reset_counters()
    resident_run(hp, dev_args, max(a.warmup, 3) * min(cps, 4), RES, a.in_flight)
    serial_kernel = serial_run(hp, dev_args, max(4, min(a.steps, 10)), RES)[1:]
    hp.span_ms()
    hp.span_ms()
    hp.reset_counters()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = hp.launch_count()
    t0 = time.perf_counter()
    per_kernel, last = resident_run(hp, dev_args, a.steps * cps, RES, a.in_flight)
    span = hp.span_ms()
    barrier()
    wall = time.perf_counter() - t0
    launches = hp.launch_count() - launches0
    out_bytes = last.out_r2_bytes + last.out_r1_bytes
    assert last.n_records == n
    allreduce_ms = None
    if world > 1:
        allreduce_ms = merged_tables_check(n * a.steps * cps * world)
    else:
        assert int(hp.counters()[1][:, 1:].sum()) == n * a.steps * cps, "counter tables do not add up to the reads processed"
    span_max = max_over_ranks(span)
    value = n * a.steps * cps * world / (span_max * 1e-3)

    progress("resident leg done")
    # ---- end to end through the C ABI: pinned host in, pinned host out; and the copies alone over the same buffers
    e2e = None
    if not a.no_e2e:
        h2 = torch.from_numpy(r2).pin_memory()
        h1 = torch.from_numpy(r1).pin_memory()
        # a chunk goes through the ABI in pieces of at most --e2e-chunk-pairs (records have a fixed size in the
        # synthetic lane, so the cuts are at record boundaries): (pointer, bytes) of both reads per piece
        assert r2.size == n * r2_rec and r1.size == n * r1_rec
        n_sub = max(1, -(-n // a.e2e_chunk_pairs))
        cuts = [n * j // n_sub for j in range(n_sub + 1)]
        sub_args = [((h2.data_ptr() + cuts[j] * r2_rec, (cuts[j + 1] - cuts[j]) * r2_rec),
                     (h1.data_ptr() + cuts[j] * r1_rec, (cuts[j + 1] - cuts[j]) * r1_rec)) for j in range(n_sub)]
        per_step_pieces = cps * n_sub

        def run_e2e(steps, flags=0, depth=2):
            d2h = 0
            pieces = steps * per_step_pieces
            done = 0
            for k in range(pieces + depth - 1):
                if k < pieces:
                    hp.submit(k, sub_args[k % n_sub][0], sub_args[k % n_sub][1], flags)
                if k >= depth - 1:
                    res = hp.collect(device_out=True)      # results are in pinned host memory; no python copy
                    d2h += res.out_r2_bytes + res.out_r1_bytes + 4 * res.unassigned.size + 4 * 8 * hp.total
                    hp.release(done)
                    done += 1
            assert done == pieces
            if world > 1:
                hp.nccl_allreduce_counters()
            return d2h // steps
        run_e2e(1)
        barrier()
        t0 = time.perf_counter()
        d2h = run_e2e(a.steps)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        # the same with MGK_OUT_GZIP: the segments leave the device as gzip members (what the reference's workers make
        # of them before they reach a file), so the D2H side carries about half the bytes
        gz_steps = max(1, min(a.steps, 4))
        # (three pieces in flight: copy in, kernels -- the deflate takes as long as a copy --, copy out)
        run_e2e(1, MGK_OUT_GZIP, 3)
        barrier()
        t0 = time.perf_counter()
        d2h_gz = run_e2e(gz_steps, MGK_OUT_GZIP, 3)
        barrier()
        dt_gz = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": n * cps * a.steps * world / dt, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes * cps),
               "d2h_bytes_per_step": int(d2h), "seconds": dt,
               "timing": "host clock between barrier+synchronize, max over ranks; pinned buffers, "
                         f"{per_step_pieces} pieces of {cuts[1]} pairs per step, 2 in flight",
               "gzip_out": {"value": n * cps * gz_steps * world / dt_gz, "unit": UNIT, "d2h_bytes_per_step": int(d2h_gz), "steps": gz_steps,
                            "what": "same call with MGK_OUT_GZIP: per-sample segments come back as gzip members made on the device; 3 pieces in flight"}}

        progress("end-to-end legs done")
        # copy-only leg: the same pinned input buffers and piece sizes, one cudaMemcpyAsync per piece and direction
        # (H2D of the text on one stream, D2H of as many bytes as the path returns on another), no kernels: what the
        # box's host<->device DMA gives this rank while every other rank does the same
        piece_out = (d2h // per_step_pieces + 4095) & ~4095
        dev_in2 = torch.empty(max(x[0][1] for x in sub_args), dtype=torch.uint8, device="cuda")
        dev_in1 = torch.empty(max(x[1][1] for x in sub_args), dtype=torch.uint8, device="cuda")
        dev_out = torch.empty(piece_out, dtype=torch.uint8, device="cuda")
        host_out = [torch.empty(piece_out, dtype=torch.uint8).pin_memory() for _ in range(2)]
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def run_copies(steps):
            for k in range(steps * per_step_pieces):
                j = k % n_sub
                with torch.cuda.stream(s_in):
                    dev_in2[:sub_args[j][0][1]].copy_(h2[cuts[j] * r2_rec:cuts[j + 1] * r2_rec], non_blocking=True)
                    dev_in1[:sub_args[j][1][1]].copy_(h1[cuts[j] * r1_rec:cuts[j + 1] * r1_rec], non_blocking=True)
                with torch.cuda.stream(s_out):
                    host_out[k & 1].copy_(dev_out, non_blocking=True)
            torch.cuda.synchronize()
        copy_steps = max(1, min(a.steps, 4))
        run_copies(1)
        barrier()
        t0 = time.perf_counter()
        run_copies(copy_steps)
        barrier()
        dtc = max_over_ranks(time.perf_counter() - t0)
        ceiling = n * cps * copy_steps * world / dtc
        e2e["copy_ceiling"] = {"value": ceiling, "unit": UNIT, "h2d_gbs": in_bytes * cps * copy_steps * world / dtc / 1e9,
                               "d2h_gbs": piece_out * per_step_pieces * copy_steps * world / dtc / 1e9, "steps": copy_steps,
                               "what": "cudaMemcpyAsync only (same pinned buffers, same piece sizes, both directions at once, all ranks at once)"}
        e2e["frac_of_copy_ceiling"] = e2e["value"] / ceiling
        del dev_in2, dev_in1, dev_out, host_out, h2, h1

    clocks = sampler.stop() if rank == 0 else None   # sampled over both timed regions (resident and end to end)

    # ---- N = 2: the product CLI with --gpus 2 (one process, two contexts, mgk_allreduce_counters) against the
    # reference's goldens, executed by rank 0 while rank 1 waits (tests/test_multi_gpu.py runs the same check)
    multi_cli = None
    if world == 2:
        if rank == 0:
            import multigpu_check
            try:
                multi_cli = {"ok": True, "files_compared": multigpu_check.cli_matches_goldens(CLI_BIN, 2)}
            except AssertionError as e:
                multi_cli = {"ok": False, "error": str(e)[:400]}
        barrier()

    def teardown():
        """Every rank leaves the same way and through normal interpreter shutdown: the library's communicator goes with
        its context, torch's process group is destroyed collectively."""
        if world > 1:
            dist.barrier()
        hp.close()
        if world > 1:
            dist.destroy_process_group()
        sys.stdout.flush()

    if rank != 0:
        teardown()
        return

    # ---- roofline of the dominant kernel (scatter + QC)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    k_ms = np.array(serial_kernel)                   # [chunks][scan, match, binscan, emit, whole], one chunk in flight
    algo_bytes = in_bytes + out_bytes                # every input byte read once, every output byte written once
    scatter_ms = float(k_ms[:, 3].mean())
    achieved = algo_bytes / (scatter_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None   # dram__bytes_read.sum + dram__bytes_write.sum of k_emit from a committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "k_emit_traffic.json")) as f:
            tr = json.load(f)
        traffic = int(tr["dram_bytes_per_launch"] * (n / tr["pairs_per_launch"]))
        traffic_src = tr.get("source", "profiles/k_emit_traffic.json") + " (ncu capture of that build, scaled to this launch; not measured in this run)"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_emit", "timing": "CUDA events on the launching stream, one chunk in flight", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo_bytes),
                "kernel_ms": {"scan_newlines": float(k_ms[:, 0].mean()), "match_plan": float(k_ms[:, 1].mean()),
                              "bin_scan": float(k_ms[:, 2].mean()), "emit_qc": scatter_ms, "chunk": float(k_ms[:, 4].mean())},
                "whole_path_frac": (algo_bytes * a.steps * cps / (span_max * 1e-3) / 1e9) / peak}

    progress("copy ceiling done")
    solo = None
    if a.gpus > 1:   # the other ranks wait in teardown()'s barrier: this GPU alone on the lane every rank just ran
        try:
            solo = short_lane_line(lane, 1 << 20, local, peak)
        except Exception as e:
            solo = {"error": f"{type(e).__name__}: {e}"[:300]}
    other = None
    if a.gpus == 1 and not a.no_other_configs:   # the BASELINE configurations that are not the headline, tracked round over round
        other = {}
        for key, ln in (("cfg3", synth.CONFIG3), ("cfg4", synth.CONFIG4), ("cfg5", synth.CONFIG5)):
            try:
                other[key] = short_lane_line(ln, 1 << 20, local, peak)
            except Exception as e:
                other[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
        try:
            other["reformat"] = short_lane_line(synth.CONFIG2, 1 << 20, local, peak, reformat=True)
        except Exception as e:
            other["reformat"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    progress("other configurations done")
    files = None
    if a.gpus == 1 and a.file_pairs > 0:
        try:
            os.sched_setaffinity(0, all_cpus)
            # in a process of its own and under a clock: the leg forks worker pools (a process that holds CUDA
            # contexts and a dozen threads is no place to fork from) and runs eight CLI commands; whatever goes wrong
            # there is reported in this line, the line itself is printed
            p = subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "file_bench.py"), str(a.file_pairs)],
                                 stdout=subprocess.PIPE, text=True, start_new_session=True)   # its progress lines go to stderr
            try:
                out, _ = p.communicate(timeout=a.file_timeout)
            except subprocess.TimeoutExpired:
                os.killpg(p.pid, signal.SIGKILL)   # the leg, its worker pools and whatever command it was running
                p.communicate()
                raise
            if p.returncode != 0:
                raise RuntimeError(f"tools/file_bench.py failed ({p.returncode}), see stderr")
            files = json.loads(out.strip().splitlines()[-1])
        except subprocess.TimeoutExpired:
            files = {"error": f"tools/file_bench.py did not finish within {a.file_timeout} s"}
        except Exception as e:
            files = {"error": f"{type(e).__name__}: {e}"[:400]}

    progress("file-level leg done")
    cpu = None
    if not a.no_cpu_baseline and a.gpus == 1:
        try:
            os.sched_setaffinity(0, all_cpus)   # the reference gets every core of the box
            cpu = cpu_run(lane, a.cpu_sample_pairs, 2, 1)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # report, never hide
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": f"failed: {e}"[:300]}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": span_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "wall_s_resident": wall, "device_span_s": span_max * 1e-3}
    if allreduce_ms is not None:
        line["allreduce"] = {"ms": allreduce_ms, "bytes": int(hp.total * (10 + hp.mism_w) * 8),
                             "checked": "reduced tables == element-wise sum of the per-rank tables, on every rank"}
    if solo is not None:
        line["single_gpu_same_lane"] = solo
    if other is not None:
        line["other_configs"] = other
    if files is not None:
        line["e2e_files"] = files
    if multi_cli is not None:
        line["cli_gpus2_parity"] = multi_cli
    print(json.dumps(line), flush=True)
    teardown()
    if multi_cli is not None and not multi_cli["ok"]:
        raise SystemExit("the product CLI with --gpus 2 does not reproduce the goldens: " + multi_cli["error"])


if __name__ == "__main__":
    main()
