"""Materialise seeded synthetic lanes as .fq.gz + sample sheet, and run a
`mgikit demultiplex`-compatible binary on them."""
from __future__ import annotations

import gzip
import hashlib
import os
import subprocess

from mgikit_b200 import synth


def write_lane(lane: synth.LaneSpec, n_pairs: int, out_dir: str, blocks: int = 1):
    os.makedirs(out_dir, exist_ok=True)
    p1 = os.path.join(out_dir, "V350012345_L01_read_1.fq.gz")
    p2 = os.path.join(out_dir, "V350012345_L01_read_2.fq.gz")
    f1 = gzip.open(p1, "wb", compresslevel=1) if lane.paired else None
    with gzip.open(p2 if lane.paired else p1, "wb", compresslevel=1) as f2:
        for b in range(blocks):
            r1, r2 = synth.make_block(lane, b, n_pairs)
            f2.write(r2.tobytes())
            if f1:
                f1.write(r1.tobytes())
    if f1:
        f1.close()
    sheet = os.path.join(out_dir, "sample_sheet.tsv")
    with open(sheet, "w") as f:
        f.write(synth.sample_sheet_text(lane))
    return (p1, p2 if lane.paired else None, sheet)


def demux_args(lane: synth.LaneSpec, r1, r2, sheet, out_dir, extra=()):
    args = ["demultiplex", "-f", r1]
    if r2:
        args += ["-r", r2]
    args += ["-s", sheet, "--template", lane.template, "-m", str(lane.mismatches), "--lane", "L01", "--run", "20240101",
             "--instrument", "INS", "-o", out_dir, "--force", "--writing-buffer-size", "4194304", "-t", "1"]
    if lane.i5_rc:
        args.append("--i5-rc")
    return args + list(extra)


def run_binary(binary, args, timeout=1800):
    p = subprocess.run([binary] + args, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, f"{binary} failed ({p.returncode}): {p.stderr[-800:]}\n{p.stdout[-400:]}"
    return p


def digest_dir(d: str) -> dict:
    """file name -> sha256 of the decompressed text (.gz) or of the raw bytes (reports)."""
    out = {}
    for name in sorted(os.listdir(d)):
        p = os.path.join(d, name)
        if not os.path.isfile(p):
            continue
        h = hashlib.sha256()
        if name.endswith(".gz"):   # streamed: gzip.decompress() is quadratic in the number of members
            if os.path.getsize(p):
                with gzip.open(p, "rb") as f:
                    for piece in iter(lambda: f.read(1 << 22), b""):
                        h.update(piece)
        else:
            with open(p, "rb") as f:
                h.update(f.read())
        out[name] = h.hexdigest()
    return out
