"""Sample list -> mgk_config, the way the C++ host does it
(mgikit_b200/csrc/host/sample_sheet.cpp: extract_templates + build_plan, which
mirror /root/reference/src/sample_manager.rs:506-730).  Used by tests and bench
to drive the C ABI without going through files."""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import List, Optional

from .capi import MGK_ABI_VERSION, MgkConfig, MgkTemplate

_RC = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def reverse_complement(s: str) -> str:
    return "".join(_RC[c] for c in reversed(s.upper()))


def parse_template(tmpl: str) -> List[int]:
    """src/sample_manager.rs:249-288"""
    g = [0] * 10
    shift = 0
    for item in reversed(tmpl.split(":")):
        kind, length = item[:2], int(item[2:])
        shift += length
        if kind == "i7":
            g[0:3] = [1, length, shift]
        elif kind == "i5":
            g[3:6] = [1, length, shift]
        elif kind == "um":
            g[6:9] = [1, length, shift]
        elif kind != "--":
            raise ValueError(f"bad template {tmpl}")
    g[9] = shift
    return g


@dataclasses.dataclass
class Sample:
    sample_id: str
    i7: str
    i5: str = "."
    template: str = "."
    i7_rc: str = "."
    i5_rc: str = "."
    job_number: str = "."


@dataclasses.dataclass
class RunSpec:
    samples: List[Sample]
    template: str = ""           # --template (overrides the sheet column)
    i7_rc: bool = False
    i5_rc: bool = False
    paired: bool = True
    read2_has_sequence: bool = True
    out_mode: int = 0            # 0 illumina, 1 mgi, 2 mgi full header
    keep_barcode: bool = False
    comprehensive_scan: bool = False
    validate: bool = False
    report_level: int = 2
    mgi_data: bool = True
    mismatches: int = 1
    per_index_error: bool = False
    l_position: int = 11
    illumina_prefix: Optional[str] = "@INS:20240101:V350012345:"
    max_chunk_bytes: int = 64 << 20
    max_chunk_records: int = 1 << 20
    n_slots: int = 2
    device: int = 0
    # `reformat` (one Sample with any index, its template list is not sent): UMI length, the sample's barcode
    reformat: bool = False
    reformat_umi_length: int = 0
    reformat_barcode: str = ""


def build_config(spec: RunSpec):
    """Returns (MgkConfig, keepalive) — keepalive owns every buffer the struct points to."""
    templates = []  # first-appearance order: dict(tmpl, has_i5, geom, i7s, i5s, rows, count)
    for s_idx, s in enumerate(spec.samples):
        tmpl = spec.template or s.template
        check_i5 = len(s.i5) > 1
        i7 = reverse_complement(s.i7) if ((spec.template and spec.i7_rc) or (not spec.template and s.i7_rc == "1")) else s.i7
        i5 = ""
        if check_i5:
            i5 = reverse_complement(s.i5) if ((spec.template and spec.i5_rc) or (not spec.template and s.i5_rc == "1")) else s.i5
        T = next((t for t in templates if t["tmpl"] == tmpl), None)
        if T is None:
            T = dict(tmpl=tmpl, has_i5=check_i5, geom=parse_template(tmpl), i7s=[], i5s=[], rows=[], count=0)
            templates.append(T)
        T["count"] += 1
        if i7 not in T["i7s"]:
            T["i7s"].append(i7)
        a = T["i7s"].index(i7)
        b = -1
        if check_i5:
            if i5 not in T["i5s"]:
                T["i5s"].append(i5)
            b = T["i5s"].index(i5)
        if any(r[0] == a and (not check_i5 or r[1] == b) for r in T["rows"]):
            raise ValueError("Two samples having the same indexes")
        T["rows"].append((a, b, s_idx))
    templates.sort(key=lambda t: -t["count"])  # stable, like the host

    keep = []
    tarr = (MgkTemplate * len(templates))()
    for k, T in enumerate(templates):
        n7, n5 = len(T["i7s"]), len(T["i5s"]) if T["has_i5"] else 0
        if T["has_i5"]:
            pairs = [-1] * (n7 * n5)
            for a, b, s_idx in T["rows"]:
                if b >= 0:
                    pairs[a * n5 + b] = s_idx
        else:
            pairs = [-1] * n7
            for a, b, s_idx in T["rows"]:
                if pairs[a] < 0:
                    pairs[a] = 1000 if b >= 0 else s_idx
        i7b = "".join(T["i7s"]).encode()
        i5b = ("".join(T["i5s"]).encode()) if T["has_i5"] else b""
        parr = (C.c_int32 * max(1, len(pairs)))(*pairs)
        keep += [i7b, i5b, parr]
        tarr[k].geometry = (C.c_int32 * 10)(*T["geom"])
        tarr[k].has_i5 = int(T["has_i5"])
        tarr[k].n_i7, tarr[k].i7 = n7, i7b
        tarr[k].n_i5, tarr[k].i5 = n5, i5b
        tarr[k].pair_sample = C.cast(parr, C.POINTER(C.c_int32))

    # get_writing_unique_samples(), src/sample_manager.rs:692-730
    first, writing = {}, []
    for i, s in enumerate(spec.samples):
        writing.append(first.setdefault(s.sample_id, i))
    n = len(spec.samples)
    writing += [n, n + 1]
    warr = (C.c_int32 * len(writing))(*writing)
    prefix = spec.illumina_prefix.encode() if (spec.illumina_prefix and spec.out_mode == 0) else None

    cfg = MgkConfig()
    cfg.abi_version = MGK_ABI_VERSION
    cfg.device = spec.device
    cfg.paired = int(spec.paired)
    cfg.read2_has_sequence = int(spec.read2_has_sequence)
    cfg.out_mode = spec.out_mode
    cfg.keep_barcode = int(spec.keep_barcode)
    cfg.comprehensive_scan = int(spec.comprehensive_scan)
    cfg.validate = int(spec.validate)
    cfg.report_level = spec.report_level
    cfg.mgi_data = int(spec.mgi_data)
    cfg.allowed_mismatches = spec.mismatches
    cfg.total_allowed_mismatches = 2 * spec.mismatches if spec.per_index_error else spec.mismatches
    cfg.l_position = spec.l_position
    cfg.illumina_prefix = prefix
    cfg.n_samples = n
    cfg.writing_sample = C.cast(warr, C.POINTER(C.c_int32))
    cfg.n_templates = len(templates)
    cfg.templates = C.cast(tarr, C.POINTER(MgkTemplate))
    cfg.max_chunk_bytes = spec.max_chunk_bytes
    cfg.max_chunk_records = spec.max_chunk_records
    cfg.n_slots = spec.n_slots
    if spec.reformat:
        bc = spec.reformat_barcode.encode()
        cfg.reformat = 1
        cfg.reformat_umi_length = spec.reformat_umi_length
        cfg.reformat_barcode = bc
        cfg.n_templates = 0
        keep.append(bc)
    keep += [tarr, warr, prefix]
    return cfg, keep
