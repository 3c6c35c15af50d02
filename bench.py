#!/usr/bin/env python
"""Benchmark of the demultiplex hot path (BASELINE.json: read pairs/s demultiplexed,
bit-exact; HBM GB/s as a fraction of peak).

A step = one pass of the whole hot path (record scan -> match -> bin scan -> scatter+QC)
over one batch of synthetic config-2 read pairs (96 samples, dual 10 bp indexes, 150 bp
pairs, mismatches=1).  `value` is measured with the batch already resident in HBM
(CUDA events on the launching stream); `e2e` goes through the public C ABI with pinned
HOST buffers, H2D and D2H copies inside the timed region.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...                     (the reference's CPU binary)
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "read_pairs_per_second_demultiplexed"
UNIT = "read pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 21, help="read pairs per step per GPU (1.5 GB of text: > 10x L2)")
    ap.add_argument("--e2e-chunk-pairs", type=int, default=1 << 20, help="end to end, a step's batch goes through the C ABI in chunks of at most this many pairs")
    ap.add_argument("--cpu-sample-pairs", type=int, default=400000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


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


def main():
    a = parse()
    from mgikit_b200 import synth
    lane = synth.CONFIG2
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    r1_rec, r2_rec = synth.record_bytes(lane)
    config = {"workload": "BASELINE configs[1]: synthetic MGI paired-end lane, 96 samples, dual 10bp indexes (i710:i510), "
                          "150bp+150bp(+20bp barcode), mismatches=1, Illumina-format output",
              "pairs_per_step_per_gpu": a.pairs, "input_bytes_per_pair": r1_rec + r2_rec,
              "l2_policy": f"each step streams {(r1_rec + r2_rec) * a.pairs >> 20} MiB in + about as much out per GPU (>> 126 MB L2); "
                           "no explicit flush needed",
              "parallelism": f"{a.gpus} GPU(s), read chunks sharded by rank, one NCCL allreduce of the counter tables"}

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
    from mgikit_b200.capi import MGK_IN_DEVICE, MGK_OUT_DEVICE, load_gpu_library
    from util import make_hotpath

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

    n = a.pairs
    r1, r2 = synth.make_block(lane, rank, n)          # every rank its own block of the lane (weak scaling)
    in_bytes = r1.size + r2.size
    spec = synth.run_spec(lane, max_chunk_bytes=max(r1.size, r2.size) + 4096, max_chunk_records=n + 64, n_slots=2, device=local)
    hp = make_hotpath(load_gpu_library(), "mgk_", spec)
    if world > 1:   # NCCL communicator of the library itself, id shared through torch.distributed
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(hp.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        hp.nccl_join(bytes(uid.cpu().numpy().tobytes()), rank, world)

    d2, d1 = torch.from_numpy(r2).cuda(), torch.from_numpy(r1).cuda()
    dev_args = ((d2.data_ptr(), d2.numel()), (d1.data_ptr(), d1.numel()))

    def run_resident(steps):
        """K steps, two chunks in flight, inputs and outputs stay in HBM."""
        per_kernel = []
        out_bytes = 0
        for k in range(steps):
            hp.submit(k, dev_args[0], dev_args[1], MGK_IN_DEVICE | MGK_OUT_DEVICE)
            if k >= 1:
                res = hp.collect(device_out=True)
                per_kernel.append(hp.last_timings())
                out_bytes = res.out_r2_bytes + res.out_r1_bytes
                hp.release(k - 1)
        res = hp.collect(device_out=True)
        per_kernel.append(hp.last_timings())
        out_bytes = res.out_r2_bytes + res.out_r1_bytes
        hp.release(steps - 1)
        if world > 1:
            hp.nccl_allreduce_counters()
        return per_kernel, out_bytes, res

    def run_serial(steps):
        """One chunk in flight at a time: per-kernel CUDA-event times that no other launch overlaps."""
        per_kernel = []
        for k in range(steps):
            hp.submit(k, dev_args[0], dev_args[1], MGK_IN_DEVICE | MGK_OUT_DEVICE)
            hp.collect(device_out=True)
            per_kernel.append(hp.last_timings())
            hp.release(k)
        return per_kernel

    # ---- resident (kernel path) timing
    hp.reset_counters()
    run_resident(max(a.warmup, 3))
    serial_kernel = run_serial(max(3, min(a.steps, 10)))[1:]
    hp.span_ms()
    hp.span_ms()
    hp.reset_counters()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = hp.launch_count()
    t0 = time.perf_counter()
    per_kernel, out_bytes, last = run_resident(a.steps)
    span = hp.span_ms()
    barrier()
    wall = time.perf_counter() - t0
    launches = hp.launch_count() - launches0
    stats, mism = hp.counters()
    total_expected = n * a.steps * (world if world > 1 else 1)
    assert int(mism[:, 1:].sum()) == total_expected, "counter tables do not add up to the reads processed"
    assert last.n_records == n
    t = torch.tensor([span], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    span_max = float(t.item())
    value = n * a.steps * world / (span_max * 1e-3)

    # ---- end to end through the C ABI: pinned host in, pinned host out
    e2e = None
    if not a.no_e2e:
        h2 = torch.from_numpy(r2).pin_memory()
        h1 = torch.from_numpy(r1).pin_memory()
        host_args = ((h2.data_ptr(), h2.numel()), (h1.data_ptr(), h1.numel()))

        # a step's batch goes through the ABI in chunks of at most --e2e-chunk-pairs (records have a fixed size in the
        # synthetic lane, so the cuts are at record boundaries): (pointer, bytes) of both reads per chunk
        assert r2.size == n * r2_rec and r1.size == n * r1_rec
        n_sub = max(1, -(-n // a.e2e_chunk_pairs))
        cuts = [n * j // n_sub for j in range(n_sub + 1)]
        sub_args = [((h2.data_ptr() + cuts[j] * r2_rec, (cuts[j + 1] - cuts[j]) * r2_rec),
                     (h1.data_ptr() + cuts[j] * r1_rec, (cuts[j + 1] - cuts[j]) * r1_rec)) for j in range(n_sub)]

        def run_e2e(steps):
            d2h = 0
            per_step = 0
            for k in range(steps * n_sub):
                hp.submit(k, sub_args[k % n_sub][0], sub_args[k % n_sub][1], 0)
                if k >= 1:
                    res = hp.collect(device_out=True)      # results are in pinned host memory; no python copy
                    per_step += res.out_r2_bytes + res.out_r1_bytes + 4 * res.unassigned.size + 4 * 8 * hp.total
                    hp.release(k - 1)
                    if k % n_sub == 0:
                        d2h, per_step = per_step, 0
            res = hp.collect(device_out=True)
            per_step += res.out_r2_bytes + res.out_r1_bytes + 4 * res.unassigned.size + 4 * 8 * hp.total
            d2h = per_step
            hp.release(steps * n_sub - 1)
            if world > 1:
                hp.nccl_allreduce_counters()
            return d2h
        run_e2e(max(2, min(a.warmup, 3)))
        barrier()
        t0 = time.perf_counter()
        d2h = run_e2e(a.steps)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": n * a.steps * world / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(in_bytes),
               "d2h_bytes_per_step": int(d2h), "timing": "host clock between barrier+synchronize, max over ranks; pinned buffers, "
                                                          f"{n_sub} chunk(s) per step, 2 chunks in flight"}

    clocks = sampler.stop() if rank == 0 else None   # sampled over both timed regions (resident and end to end)
    def teardown():
        """Every rank leaves the same way: NCCL communicators (the library's and torch's) are destroyed collectively;
        a watchdog ends the process if a peer is already gone."""
        import threading
        threading.Timer(30.0, lambda: os._exit(0)).start()
        if world > 1:
            dist.barrier()
        hp.close()
        if world > 1:
            dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(0)

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
    k_ms = np.array(serial_kernel)                   # [steps][scan, match, binscan, emit, whole], one chunk in flight
    algo_bytes = in_bytes + out_bytes                # every input byte read once, every output byte written once
    scatter_ms = float(k_ms[:, 3].mean())
    achieved = algo_bytes / (scatter_ms * 1e-3) / 1e9
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of k_emit from the committed ncu capture, scaled per pair
    try:
        with open(os.path.join(ROOT, "profiles", "k_emit_traffic.json")) as f:
            tr = json.load(f)
        traffic = int(tr["dram_bytes_per_launch"] * (n / tr["pairs_per_launch"]))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_emit", "timing": "CUDA events on the launching stream, one chunk in flight", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo_bytes),
                "kernel_ms": {"scan_newlines": float(k_ms[:, 0].mean()), "match_plan": float(k_ms[:, 1].mean()),
                              "bin_scan": float(k_ms[:, 2].mean()), "emit_qc": scatter_ms, "chunk": float(k_ms[:, 4].mean())},
                "whole_path_frac": (algo_bytes * a.steps / (span_max * 1e-3) / 1e9) / peak}
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
            "roofline": roofline, "cpu_baseline": cpu, "wall_s_resident": wall}
    print(json.dumps(line), flush=True)
    teardown()


if __name__ == "__main__":
    main()
