"""ctypes mirror of include/mgikit_gpu.h.

``HotPath`` wraps ONE implementation of the C ABI, selected by (library, symbol
prefix): the product (``libmgikit_gpu.so``, ``mgk_``) or — in tests only — the
oracle (``orc_``) / the CPU emulation of the device logic (``emu_``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MGK_ABI_VERSION = 2
MGK_IN_DEVICE = 1
MGK_OUT_DEVICE = 2
MGK_OUT_GZIP = 4
MGK_N_STATS = 10
MGK_NCCL_ID_BYTES = 128
PKG_DIR = os.path.dirname(os.path.abspath(__file__))
GPU_LIB = os.path.join(PKG_DIR, "lib", "libmgikit_gpu.so")
CLI_BIN = os.path.join(PKG_DIR, "bin", "mgikit")


class MgkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class MgkTemplate(C.Structure):
    _fields_ = [("geometry", C.c_int32 * 10), ("has_i5", C.c_int32), ("n_i7", C.c_int32), ("i7", C.c_char_p),
                ("n_i5", C.c_int32), ("i5", C.c_char_p), ("pair_sample", C.POINTER(C.c_int32))]


class MgkConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("paired", C.c_int32),
                ("read2_has_sequence", C.c_int32), ("out_mode", C.c_int32), ("keep_barcode", C.c_int32),
                ("comprehensive_scan", C.c_int32), ("validate", C.c_int32), ("report_level", C.c_int32),
                ("mgi_data", C.c_int32), ("allowed_mismatches", C.c_int32), ("total_allowed_mismatches", C.c_int32),
                ("l_position", C.c_int32), ("illumina_prefix", C.c_char_p), ("n_samples", C.c_int32),
                ("writing_sample", C.POINTER(C.c_int32)), ("n_templates", C.c_int32),
                ("templates", C.POINTER(MgkTemplate)), ("max_chunk_bytes", C.c_uint64),
                ("max_chunk_records", C.c_uint64), ("n_slots", C.c_int32),
                # `mgikit reformat`: one sample, the barcode is the one in the header
                ("reformat", C.c_int32), ("reformat_umi_length", C.c_int32), ("reformat_barcode", C.c_char_p)]


class MgkResult(C.Structure):
    _fields_ = [("seq", C.c_uint64), ("n_records", C.c_uint64), ("consumed_r2", C.c_uint64),
                ("consumed_r1", C.c_uint64), ("out_r2", C.c_void_p), ("out_r1", C.c_void_p),
                ("out_r2_bytes", C.c_uint64), ("out_r1_bytes", C.c_uint64), ("seg_off_r2", C.POINTER(C.c_uint64)),
                ("seg_len_r2", C.POINTER(C.c_uint64)), ("seg_off_r1", C.POINTER(C.c_uint64)),
                ("seg_len_r1", C.POINTER(C.c_uint64)), ("unassigned", C.POINTER(C.c_uint32)),
                ("n_unassigned", C.c_uint64), ("validate_error", C.c_int32), ("validate_record", C.c_uint64)]


# every entry point include/mgikit_gpu.h declares
ABI_SYMBOLS = ["create", "destroy", "last_error", "submit", "collect", "release", "counters", "reset_counters",
               "allreduce_counters", "reduced_counters", "nccl_unique_id", "nccl_join", "nccl_allreduce_counters", "host_alloc",
               "host_free", "last_timings", "last_gzip_ms", "span_ms", "launch_count", "scan_newlines", "match_barcodes"]


def load_gpu_library() -> C.CDLL:
    """The product library.  Raises if it was not built — there is no fallback."""
    if not os.path.exists(GPU_LIB):
        raise MgkError(-2, f"{GPU_LIB} is missing: run `make` (or __graft_entry__.build()); there is no CPU fallback")
    return C.CDLL(GPU_LIB, mode=C.RTLD_GLOBAL)


class Result:
    """One chunk's result copied out of the library-owned buffers."""

    def __init__(self, res: MgkResult, total: int, device_out: bool = False):
        self.seq = res.seq
        self.n_records = res.n_records
        self.consumed_r2, self.consumed_r1 = res.consumed_r2, res.consumed_r1
        self.out_r2_bytes, self.out_r1_bytes = res.out_r2_bytes, res.out_r1_bytes
        self.seg_off_r2 = np.array(res.seg_off_r2[:total], dtype=np.uint64)
        self.seg_len_r2 = np.array(res.seg_len_r2[:total], dtype=np.uint64)
        self.seg_off_r1 = np.array(res.seg_off_r1[:total], dtype=np.uint64)
        self.seg_len_r1 = np.array(res.seg_len_r1[:total], dtype=np.uint64)
        self.validate_error, self.validate_record = res.validate_error, res.validate_record
        self.out_r2_ptr, self.out_r1_ptr = res.out_r2, res.out_r1
        if device_out:
            self.out_r2 = self.out_r1 = None
        else:
            self.out_r2 = C.string_at(res.out_r2, res.out_r2_bytes) if res.out_r2_bytes else b""
            self.out_r1 = C.string_at(res.out_r1, res.out_r1_bytes) if (res.out_r1 and res.out_r1_bytes) else b""
        n = res.n_unassigned
        self.unassigned = np.ctypeslib.as_array(res.unassigned, shape=(n,)).copy() if n else np.zeros(0, np.uint32)

    def segment(self, which: int, b: int) -> bytes:
        buf, off, ln = (self.out_r2, self.seg_off_r2, self.seg_len_r2) if which == 2 else \
            (self.out_r1, self.seg_off_r1, self.seg_len_r1)
        return buf[int(off[b]):int(off[b]) + int(ln[b])]


class HotPath:
    def __init__(self, lib: C.CDLL, prefix: str, cfg: MgkConfig, keepalive=None):
        self.lib, self.prefix, self.cfg, self._keep = lib, prefix, cfg, keepalive
        self.total = cfg.n_samples + 2
        self.mism_w = 2 * cfg.allowed_mismatches + 2
        f = self._f
        f("create").argtypes = [C.POINTER(MgkConfig), C.POINTER(C.c_void_p)]
        f("destroy").argtypes = [C.c_void_p]
        f("destroy").restype = None
        f("last_error").argtypes = [C.c_void_p]
        f("last_error").restype = C.c_char_p
        f("submit").argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32]
        f("collect").argtypes = [C.c_void_p, C.POINTER(MgkResult)]
        f("release").argtypes = [C.c_void_p, C.c_uint64]
        f("counters").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.ctx = C.c_void_p()
        rc = f("create")(C.byref(cfg), C.byref(self.ctx))
        if rc != 0:
            raise MgkError(rc, (f("last_error")(None) or b"").decode())

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, rc):
        if rc != 0:
            raise MgkError(rc, (self._f("last_error")(self.ctx) or b"").decode())

    def close(self):
        if self.ctx:
            self._f("destroy")(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _ptr(buf):
        if buf is None:
            return None, 0
        if isinstance(buf, tuple):          # (device pointer, length)
            return C.c_void_p(buf[0]), buf[1]
        if isinstance(buf, (bytes, bytearray)):
            arr = np.frombuffer(buf, dtype=np.uint8)
        else:
            arr = np.ascontiguousarray(buf, dtype=np.uint8)
        return C.c_void_p(arr.ctypes.data), arr.size

    def submit(self, seq: int, r2, r1=None, flags: int = 0):
        p2, n2 = self._ptr(r2)
        p1, n1 = self._ptr(r1)
        self._hold = (r2, r1)
        self._check(self._f("submit")(self.ctx, seq, p2, n2, p1, n1, flags))

    def collect(self, device_out: bool = False) -> Result:
        res = MgkResult()
        self._check(self._f("collect")(self.ctx, C.byref(res)))
        return Result(res, self.total, device_out)

    def release(self, seq: int):
        self._check(self._f("release")(self.ctx, seq))

    def process(self, r2, r1=None, seq: int = 0, flags: int = 0) -> Result:
        self.submit(seq, r2, r1, flags)
        out = self.collect()
        self.release(seq)
        return out

    def counters(self):
        stats = np.zeros((self.total, MGK_N_STATS), dtype=np.uint64)
        mism = np.zeros((self.total, self.mism_w), dtype=np.uint64)
        self._check(self._f("counters")(self.ctx, stats.ctypes.data, mism.ctypes.data))
        return stats, mism

    def reduced_counters(self):
        """Sum over all ranks as left by the last counter all-reduce (the running tables stay local)."""
        stats = np.zeros((self.total, MGK_N_STATS), dtype=np.uint64)
        mism = np.zeros((self.total, self.mism_w), dtype=np.uint64)
        fn = self._f("reduced_counters")
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(fn(self.ctx, stats.ctypes.data, mism.ctypes.data))
        return stats, mism

    def reset_counters(self):
        fn = self._f("reset_counters")
        fn.argtypes = [C.c_void_p]
        self._check(fn(self.ctx))

    def scan_newlines(self, buf, capacity=None):
        p, n = self._ptr(buf)
        cap = capacity if capacity is not None else max(16, n)
        pos = np.zeros(cap, dtype=np.uint32)
        cnt = C.c_uint64()
        fn = self._f("scan_newlines")
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        self._check(fn(self.ctx, p, n, pos.ctypes.data, cap, C.byref(cnt)))
        return pos[:cnt.value]

    def match_barcodes(self, barcodes: np.ndarray):
        arr = np.ascontiguousarray(barcodes, dtype=np.uint8)
        n = arr.shape[0]
        s = np.zeros(n, dtype=np.int32)
        m = np.zeros(n, dtype=np.int32)
        fn = self._f("match_barcodes")
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        self._check(fn(self.ctx, arr.ctypes.data, n, s.ctypes.data, m.ctypes.data))
        return s, m

    def last_timings(self):
        ms = (C.c_float * 5)()
        fn = self._f("last_timings")
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        self._check(fn(self.ctx, ms))
        return list(ms)

    def last_gzip_ms(self) -> float:
        ms = C.c_float()
        fn = self._f("last_gzip_ms")
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        self._check(fn(self.ctx, C.byref(ms)))
        return float(ms.value)

    def span_ms(self) -> float:
        ms = C.c_float()
        fn = self._f("span_ms")
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        self._check(fn(self.ctx, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        fn = self._f("launch_count")
        fn.argtypes = [C.c_void_p]
        fn.restype = C.c_uint64
        return int(fn(self.ctx))

    # NCCL merge of the counter tables, one process per GPU
    def nccl_unique_id(self) -> bytes:
        buf = (C.c_uint8 * MGK_NCCL_ID_BYTES)()
        fn = self._f("nccl_unique_id")
        fn.argtypes = [C.c_void_p]
        rc = fn(buf)
        if rc != 0:
            raise MgkError(rc, (self._f("last_error")(None) or b"").decode())
        return bytes(buf)

    def nccl_join(self, uid: bytes, rank: int, nranks: int):
        buf = (C.c_uint8 * MGK_NCCL_ID_BYTES).from_buffer_copy(uid)
        fn = self._f("nccl_join")
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        self._check(fn(self.ctx, buf, rank, nranks))

    def nccl_allreduce_counters(self):
        fn = self._f("nccl_allreduce_counters")
        fn.argtypes = [C.c_void_p]
        self._check(fn(self.ctx))
