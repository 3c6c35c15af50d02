#!/usr/bin/env python
"""File-level end to end (bullet 1): `.fq.gz` in -> per-sample `.fastq.gz` + reports out,
the drop-in CLI next to the reference's own binary ON THE SAME FILES, same box, thread counts stated.

The lane is block-streamed (block b is a pure function of (seed, b), synth.make_block) and written as one gzip
member per block by a pool of processes -- the multi-member layout of bgzip / pigz -i / concatenated lanes /
mgikit's own output, which both binaries read (the reference through flate2's MultiGzDecoder).  Parity inside
the bench: every report file of the full run byte-identical to the reference's, and every FASTQ (decompressed)
of a prefix lane identical to the reference at -t 1 (its multi-threaded FASTQ order is arbitrary).

  python tools/file_bench.py [pairs]            prints the JSON block bench.py embeds as `e2e_files`
"""
from __future__ import annotations

import gzip
import hashlib
import json
import multiprocessing as mp
import os
import shutil
import subprocess
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BLOCK_PAIRS = 250_000


def _gz_member(raw: bytes) -> bytes:
    c = zlib.compressobj(1, zlib.DEFLATED, 31)
    return c.compress(raw) + c.flush()


def _block(args):
    lane, b, n, first = args
    from mgikit_b200 import synth
    r1, r2 = synth.make_block(lane, b, n, first_read=first)
    return _gz_member(r1.tobytes()) if r1 is not None else b"", _gz_member(r2.tobytes())


def _rf_block(args):
    lane, b, n = args
    from mgikit_b200 import synth
    r1, r2 = synth.reformat_block(lane, b, n)
    return _gz_member(r1.tobytes()), _gz_member(r2.tobytes())


def write_reformat_lane(lane, n_pairs, out_dir, block_pairs=BLOCK_PAIRS):
    """One sample's reads with ` n:N:0:BARCODE` headers (synth.reformat_block), one gzip member per block, named the way
    `reformat` reads sample label and lane from (Flowcell_Lane_Label_n.fq.gz)."""
    os.makedirs(out_dir, exist_ok=True)
    p1 = os.path.join(out_dir, "V350012345_L01_S001_1.fq.gz")
    p2 = os.path.join(out_dir, "V350012345_L01_S001_2.fq.gz")
    jobs = [(lane, b, min(block_pairs, n_pairs - b * block_pairs)) for b in range(-(-n_pairs // block_pairs))]
    with mp.get_context("fork").Pool(max(1, min(len(jobs), os.cpu_count() or 2))) as pool, open(p1, "wb") as f1, open(p2, "wb") as f2:
        for z1, z2 in pool.imap(_rf_block, jobs):
            f1.write(z1)
            f2.write(z2)
    return p1, p2


def write_lane_blocks(lane, n_pairs, out_dir, block_pairs=BLOCK_PAIRS, single_member=False):
    """The lane as .fq.gz, one gzip member per block of `block_pairs` (or one member in all, python's gzip
    writer, when single_member), read numbers continuing across blocks; returns (r1, r2, sheet, seconds)."""
    from mgikit_b200 import synth
    os.makedirs(out_dir, exist_ok=True)
    t0 = time.perf_counter()
    p1 = os.path.join(out_dir, "V350012345_L01_read_1.fq.gz")
    p2 = os.path.join(out_dir, "V350012345_L01_read_2.fq.gz")
    jobs, first = [], 0
    while first < n_pairs:
        n = min(block_pairs, n_pairs - first)
        jobs.append((lane, len(jobs), n, first))
        first += n
    if single_member:
        f1 = gzip.open(p1, "wb", compresslevel=1) if lane.paired else None
        with gzip.open(p2 if lane.paired else p1, "wb", compresslevel=1) as f2:
            for j in jobs:
                r1, r2 = synth.make_block(j[0], j[1], j[2], first_read=j[3])
                f2.write(r2.tobytes())
                if f1:
                    f1.write(r1.tobytes())
        if f1:
            f1.close()
    else:
        procs = max(1, min(len(jobs), (os.cpu_count() or 2)))
        with mp.get_context("fork").Pool(procs) as pool, open(p2 if lane.paired else p1, "wb") as f2:
            f1 = open(p1, "wb") if lane.paired else None
            for z1, z2 in pool.imap(_block, jobs):
                f2.write(z2)
                if f1:
                    f1.write(z1)
            if f1:
                f1.close()
    sheet = os.path.join(out_dir, "sample_sheet.tsv")
    with open(sheet, "w") as f:
        f.write(synth.sample_sheet_text(lane))
    return p1, (p2 if lane.paired else None), sheet, time.perf_counter() - t0


def _scratch(need_bytes):
    """tmpfs when it has room (the disk of a GPU box is not what is being measured), else the default temp dir."""
    for base in ("/dev/shm", tempfile.gettempdir()):
        try:
            st = os.statvfs(base)
            if st.f_bavail * st.f_frsize > need_bytes:
                return base, st.f_bavail * st.f_frsize
        except OSError:
            pass
    st = os.statvfs(tempfile.gettempdir())
    return tempfile.gettempdir(), st.f_bavail * st.f_frsize


def _timed(binary, args, timeout=240, want_log=False):
    print(f"[file_bench] {os.path.basename(binary)} {args[0]} -> {args[args.index('-o') + 1] if '-o' in args else ''}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    p = subprocess.run([binary] + args, capture_output=True, text=True, timeout=timeout)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{os.path.basename(binary)} failed ({p.returncode}): {p.stderr[-500:]}")
    return (dt, p.stdout) if want_log else dt


def _phases(log):
    """The CLI's own account of a run: set-up (driver + contexts + pinned buffers), the pump phase, where the pump waited."""
    import re
    out = {}
    m = re.search(r"hot path contexts ready after ([0-9.]+) s", log)
    if m:
        out["setup_seconds"] = float(m.group(1))
    m = re.search(r"pump phase: (\d+) chunks in ([0-9.]+) s", log)
    if m:
        out["pump_seconds"] = float(m.group(2))
        out["chunks"] = int(m.group(1))
    m = re.search(r"pump waited ([0-9.]+) s for input, ([0-9.]+) s for the device, ([0-9.]+) s for the output side, spent ([0-9.]+) s submitting", log)
    if m:
        out["pump_waits_seconds"] = {"input": float(m.group(1)), "device": float(m.group(2)), "output": float(m.group(3)), "submit": float(m.group(4))}
    return out


def _gz_digest(path):
    """sha256 of the decompressed text, all members like MultiGzDecoder -- streamed: gzip.decompress() copies the rest of
    the file once per member, hours for the 23 000 members of a one-sample output."""
    h = hashlib.sha256()
    if os.path.getsize(path):
        with gzip.open(path, "rb") as f:
            for piece in iter(lambda: f.read(1 << 22), b""):
                h.update(piece)
    return h.hexdigest()


def _digest_dir(d, gz):
    out = {}
    names = sorted(n for n in os.listdir(d) if os.path.isfile(os.path.join(d, n)) and n.endswith(".gz") == gz)
    if gz:
        with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
            for n, h in zip(names, pool.map_async(_gz_digest, [os.path.join(d, n) for n in names]).get(timeout=300)):
                out[n] = h
    else:
        for n in names:
            with open(os.path.join(d, n), "rb") as f:
                out[n] = hashlib.sha256(f.read()).hexdigest()
    return out


def run(lane, n_pairs, product_cli, ref_bin, prefix_pairs=1_000_000, side_figures=True):
    import lanes
    cores = os.cpu_count() or 1
    r1_rec = 336 if lane.paired else 0
    need = int(n_pairs * (r1_rec + 376) * 0.55 * 3.2) + (2 << 30)    # gz input + two sets of gz output + slack
    base, free = _scratch(need)
    if free < need:   # state it, never fail silently: shrink the lane to what the box holds
        n_pairs = max(prefix_pairs, int(n_pairs * free / need * 0.8))
    d = tempfile.mkdtemp(prefix="mgk_files_", dir=base)
    out = {"lane": lane.name, "pairs": n_pairs, "scratch": base, "host_cores": cores,
           "layout": f"one gzip member per block of {BLOCK_PAIRS} pairs (gzip -1), both reads"}
    try:
        r1, r2, sheet, gen_s = write_lane_blocks(lane, n_pairs, os.path.join(d, "in"))
        out["generate_seconds"] = round(gen_s, 2)
        out["input_gz_bytes"] = os.path.getsize(r1) + (os.path.getsize(r2) if r2 else 0)

        def args_for(out_dir, threads, extra=()):
            a = lanes.demux_args(lane, r1, r2, sheet, out_dir, list(extra))
            a[a.index("-t") + 1] = str(threads)
            return a
        # ---- the drop-in CLI, twice (the second run has the page cache the reference run will have too)
        ours_dir = os.path.join(d, "ours")
        _timed(product_cli, args_for(ours_dir, cores))
        t_ours, log = _timed(product_cli, args_for(ours_dir, cores), want_log=True)
        out["ours"] = {"pairs_per_s": n_pairs / t_ours, "seconds": t_ours, "threads": cores, "cmd": "mgikit_b200/bin/mgikit demultiplex -t %d" % cores}
        ph = _phases(log)
        if "pump_seconds" in ph:   # the run minus process set-up and exit: what a full flow cell would see
            ph["pump_pairs_per_s"] = n_pairs / ph["pump_seconds"]
        out["ours"]["phases"] = ph
        # ---- the reference binary on the same files
        ref_dir = os.path.join(d, "ref")
        t_ref = _timed(ref_bin, args_for(ref_dir, cores))
        out["reference"] = {"pairs_per_s": n_pairs / t_ref, "seconds": t_ref, "threads": cores,
                            "cmd": "mgikit demultiplex -t %d --writing-buffer-size 4194304" % cores}
        out["ratio"] = t_ref / t_ours
        # ---- parity, full run: every report byte for byte
        a, b = _digest_dir(ours_dir, gz=False), _digest_dir(ref_dir, gz=False)
        if a != b:
            raise AssertionError(f"reports of the full run differ from the reference's: {sorted(k for k in set(a) | set(b) if a.get(k) != b.get(k))[:5]}")
        gz_ours = sorted(n for n in os.listdir(ours_dir) if n.endswith(".gz"))
        gz_ref = sorted(n for n in os.listdir(ref_dir) if n.endswith(".gz"))
        if gz_ours != gz_ref:
            raise AssertionError("the full run wrote a different set of FASTQ files than the reference")
        out["parity"] = {"reports_full_run": f"{len(a)} report files byte-identical", "fastq_files_full_run": len(gz_ours)}
        out["output_gz_bytes"] = {"ours": sum(os.path.getsize(os.path.join(ours_dir, n)) for n in gz_ours),
                                  "reference": sum(os.path.getsize(os.path.join(ref_dir, n)) for n in gz_ref)}
        shutil.rmtree(ref_dir, ignore_errors=True)
        shutil.rmtree(ours_dir, ignore_errors=True)
        if side_figures:   # BASELINE.md §3 side figures + where our own time goes
            t = _timed(product_cli, args_for(ours_dir, cores, ["--host-gzip"]))
            out["ours_host_gzip"] = {"pairs_per_s": n_pairs / t, "seconds": t, "what": "gzip members made by zlib on the host threads instead of on the device"}
            shutil.rmtree(ours_dir, ignore_errors=True)
            t = _timed(ref_bin, args_for(ref_dir, cores, ["--compression-level", "0", "--report-level", "0"]))
            out["reference_no_codec_out"] = {"pairs_per_s": n_pairs / t, "seconds": t, "what": "--compression-level 0 --report-level 0"}
            shutil.rmtree(ref_dir, ignore_errors=True)
        # ---- parity, prefix lane: every FASTQ (decompressed) and report equal to the reference at -t 1
        pre = min(prefix_pairs, n_pairs)
        q1, q2, qsheet, _ = write_lane_blocks(lane, pre, os.path.join(d, "pre"))

        def pre_args(out_dir, threads):
            a = lanes.demux_args(lane, q1, q2, qsheet, out_dir)
            a[a.index("-t") + 1] = str(threads)
            return a
        _timed(product_cli, pre_args(ours_dir, cores))
        t1 = _timed(ref_bin, pre_args(ref_dir, 1))
        out["reference_t1"] = {"pairs_per_s": pre / t1, "seconds": t1, "pairs": pre, "threads": 1}
        for gz in (False, True):
            a, b = _digest_dir(ours_dir, gz), _digest_dir(ref_dir, gz)
            if a != b:
                raise AssertionError(f"prefix lane: {'FASTQ' if gz else 'report'} files differ from the reference at -t 1: "
                                     f"{sorted(k for k in set(a) | set(b) if a.get(k) != b.get(k))[:5]}")
            out["parity"]["fastq_prefix" if gz else "reports_prefix"] = f"{len(a)} files identical on a {pre}-pair prefix vs the reference at -t 1"
        shutil.rmtree(ref_dir, ignore_errors=True)
        shutil.rmtree(ours_dir, ignore_errors=True)
        if side_figures:
            try:
                _single_member_figures(out, lane, n_pairs, d, ours_dir, product_cli, ref_bin, cores)
            except Exception as e:
                out["ours_single_member_input"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            try:
                _reformat_figures(out, lane, n_pairs, d, ours_dir, ref_dir, product_cli, ref_bin, cores)
            except Exception as e:
                out["reformat"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _single_member_figures(out, lane, n_pairs, d, ours_dir, product_cli, ref_bin, cores):
    """One member per file (what a plain `gzip` writes): each member is decoded in parts by the file's inflate threads."""
    import lanes
    one = min(n_pairs, 2_000_000)
    s1, s2, ssheet, _ = write_lane_blocks(lane, one, os.path.join(d, "single"), single_member=True)
    a = lanes.demux_args(lane, s1, s2, ssheet, ours_dir)
    a[a.index("-t") + 1] = str(cores)
    _timed(product_cli, a)
    t, log = _timed(product_cli, a, want_log=True)
    out["ours_single_member_input"] = {"pairs_per_s": one / t, "seconds": t, "pairs": one, "phases": _phases(log),
                                       "what": "the same lane as ONE gzip member per file (what plain gzip writes): each member is decoded in parts by all the inflate threads of its file"}
    t = _timed(ref_bin, a)
    out["reference_single_member_input"] = {"pairs_per_s": one / t, "seconds": t, "pairs": one, "threads": cores}
    shutil.rmtree(ours_dir, ignore_errors=True)   # (both runs wrote there)


def _reformat_figures(out, lane, n_pairs, d, ours_dir, ref_dir, product_cli, ref_bin, cores):
    """`reformat`: one sample's reads, Illumina headers + UMI into the header + quality report, both binaries."""
    from mgikit_b200 import synth
    nrf = min(n_pairs, 2_000_000)
    shutil.rmtree(ours_dir, ignore_errors=True)
    shutil.rmtree(ref_dir, ignore_errors=True)
    f1, f2 = write_reformat_lane(lane, nrf, os.path.join(d, "rf"))
    a = ["reformat", "-f", f1, "-r", f2, "--lane", "L01", "--run", "20240101", "--instrument", "INS", "--force",
         "--umi-length", "8", "--barcode", synth.REFORMAT_BARCODE, "-o"]
    _timed(product_cli, a + [ours_dir, "-t", str(cores)])
    t = _timed(product_cli, a + [ours_dir, "-t", str(cores)])
    t_r = _timed(ref_bin, a + [ref_dir])
    for gz in (False, True):
        x, y = _digest_dir(ours_dir, gz), _digest_dir(ref_dir, gz)
        if x != y or not x:
            raise AssertionError(f"reformat: {'FASTQ' if gz else 'report'} files differ from the reference's")
    out["reformat"] = {"pairs": nrf, "ours": {"pairs_per_s": nrf / t, "seconds": t, "threads": cores},
                       "reference": {"pairs_per_s": nrf / t_r, "seconds": t_r, "threads": "its own (the command has no thread option)"},
                       "ratio": t_r / t, "parity": "both FASTQ files (decompressed) and the sample_stats report identical",
                       "cmd": "mgikit reformat --umi-length 8 --barcode " + synth.REFORMAT_BARCODE}
    shutil.rmtree(ref_dir, ignore_errors=True)
    shutil.rmtree(ours_dir, ignore_errors=True)


if __name__ == "__main__":
    from mgikit_b200 import synth
    from mgikit_b200.capi import CLI_BIN
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    ref = os.path.join(ROOT, "oracle", "_ref", "mgikit")
    print(json.dumps(run(synth.CONFIG2, n, CLI_BIN, ref)))
