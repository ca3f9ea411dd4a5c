"""ctypes binding of libfeht_cuda.so (include/feht_cuda.h) -- the Python test/bench harness.

This is a thin 1:1 wrapper: every method is one C-ABI call with numpy host buffers.  There is no
Python or CPU fallback: if the shared library is missing or no B200 is visible the import of the
library / `Context()` raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / "libfeht_cuda.so"

OK = 0
CORRECTION_NONE, CORRECTION_BONFERRONI = 0, 1
SORT_ABS_RATIO_DESC, SORT_PVALUE_ASC = 0, 1
ENGINE_AUTO, ENGINE_POPC, ENGINE_GEMM, ENGINE_GEMM_FP4, ENGINE_GEMM_F8 = 0, 1, 2, 3, 4
FET_TWO_TAIL, FET_ONE_TAIL = 0, 1
E_INVALID, E_CUDA, E_BOUNDS, E_NOMEM, E_NODEVICE = -1, -2, -3, -4, -5

# every symbol include/feht_cuda.h declares (tests check the library exports all of them)
EXPORTS = [
    "feht_create", "feht_destroy", "feht_last_error", "feht_set_stream", "feht_version",
    "feht_set_samples", "feht_reserve_rows", "feht_clear_rows", "feht_add_rows_chars",
    "feht_add_rows_packed", "feht_add_snp_packed", "feht_add_snp_chars", "feht_generate_synthetic",
    "feht_read_plane", "feht_row_stride_words", "feht_num_rows", "feht_add_comparison",
    "feht_add_category", "feht_clear_comparisons", "feht_num_comparisons", "feht_comparison_info",
    "feht_set_count_engine", "feht_run", "feht_result_count", "feht_result_total",
    "feht_result_fetch", "feht_set_row_base", "feht_last_timings", "feht_merge_sorted",
    "feht_choose_table", "feht_eval_tables", "feht_result_fetch_device", "feht_merge_sorted_device",
    "feht_scatter_rows_device", "feht_run_async", "feht_timings_accumulated",
    "feht_set_host_threads", "feht_last_upload_host_rows", "feht_add_rows_text", "feht_text_lines", "feht_set_name_rank", "feht_set_fet_mode", "feht_pack_chars_host",
    "feht_create_multi", "feht_num_devices", "feht_merge_shards_device", "feht_set_comparison_window",
    # include/feht_host.h
    "feht_host_run", "feht_host_plan", "feht_host_show_double", "feht_host_free",
]


class FehtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"feht error {code}: {msg}")
        self.code = code
        self.msg = msg


class SynthSpec(C.Structure):
    """feht_synth_spec (include/feht_cuda.h) == fo_synth_spec (oracle/feht_oracle.h)."""
    _fields_ = [
        ("seed", C.c_uint64),
        ("n_samples", C.c_int64),
        ("missing_thr", C.c_uint32),
        ("planted_per_million", C.c_uint32),
        ("split_col", C.c_int64),
        ("thr_hi", C.c_uint32),
        ("thr_lo", C.c_uint32),
        ("freq_thr", C.c_uint32 * 256),
    ]


class ShardRows(C.Structure):
    """feht_shard_rows: device columns of one shard's ordered rows (all comparisons), host segment offsets."""
    _fields_ = [("row", C.c_void_p), ("goa", C.c_void_p), ("gob", C.c_void_p), ("gta", C.c_void_p),
                ("gtb", C.c_void_p), ("p", C.c_void_p), ("ratio", C.c_void_p), ("name_rank", C.c_void_p),
                ("seg_off", C.c_void_p)]


_lib = None


def lib() -> C.CDLL:
    """Load libfeht_cuda.so; loud failure when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m feht_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(str(_LIB_PATH))
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    L.feht_create.argtypes = [P(vp), C.c_int]
    L.feht_create.restype = C.c_int
    L.feht_create_multi.argtypes = [P(vp), P(C.c_int), C.c_int]
    L.feht_create_multi.restype = C.c_int
    L.feht_num_devices.argtypes = [vp]
    L.feht_merge_shards_device.argtypes = [vp, C.c_int, i64, P(ShardRows), C.c_int, P(ShardRows)]
    L.feht_destroy.argtypes = [vp]
    L.feht_destroy.restype = None
    L.feht_last_error.argtypes = [vp]
    L.feht_last_error.restype = C.c_char_p
    L.feht_set_stream.argtypes = [vp, vp]
    L.feht_version.argtypes = []
    L.feht_version.restype = C.c_char_p
    L.feht_set_samples.argtypes = [vp, i64]
    L.feht_reserve_rows.argtypes = [vp, i64]
    L.feht_clear_rows.argtypes = [vp]
    L.feht_add_rows_chars.argtypes = [vp, vp, vp, i64, vp]
    L.feht_add_snp_chars.argtypes = [vp, vp, vp, i64, vp]
    L.feht_add_rows_packed.argtypes = [vp, vp, vp, i64, i64, vp]
    L.feht_add_snp_packed.argtypes = [vp, vp, vp, vp, i64, i64, vp]
    L.feht_generate_synthetic.argtypes = [vp, P(SynthSpec), C.c_int, i64, i64]
    L.feht_read_plane.argtypes = [vp, C.c_int, i64, i64, vp]
    L.feht_row_stride_words.argtypes = [vp]
    L.feht_row_stride_words.restype = i64
    L.feht_num_rows.argtypes = [vp]
    L.feht_num_rows.restype = i64
    L.feht_add_comparison.argtypes = [vp, vp, i64, vp, i64]
    L.feht_add_comparison.restype = i64
    L.feht_add_category.argtypes = [vp, vp, i32]
    L.feht_add_category.restype = i64
    L.feht_clear_comparisons.argtypes = [vp]
    L.feht_num_comparisons.argtypes = [vp]
    L.feht_num_comparisons.restype = i64
    L.feht_comparison_info.argtypes = [vp, i64, P(i32), P(i32), P(i32), P(i32)]
    L.feht_set_count_engine.argtypes = [vp, C.c_int]
    L.feht_run.argtypes = [vp, C.c_int, f64, C.c_int, i64]
    L.feht_set_host_threads.argtypes = [vp, C.c_int]
    L.feht_pack_chars_host.argtypes = [vp, vp, i64, i64, C.c_int, i64, vp, vp, vp, C.c_int]
    L.feht_set_fet_mode.argtypes = [vp, C.c_int]
    L.feht_last_upload_host_rows.argtypes = [vp]
    L.feht_last_upload_host_rows.restype = i64
    L.feht_add_rows_text.argtypes = [vp, vp, i64, C.c_char, C.c_int, vp]
    L.feht_add_rows_text.restype = C.c_int
    L.feht_text_lines.argtypes = [vp, vp, vp, i64]
    L.feht_text_lines.restype = C.c_int
    L.feht_set_name_rank.argtypes = [vp, vp, i64]
    L.feht_set_name_rank.restype = C.c_int
    L.feht_run_async.argtypes = [vp, C.c_int, f64, C.c_int, i64]
    L.feht_timings_accumulated.argtypes = [vp, P(f64), C.c_int, C.c_int]
    L.feht_result_count.argtypes = [vp, i64]
    L.feht_result_count.restype = i64
    L.feht_result_total.argtypes = [vp]
    L.feht_result_total.restype = i64
    L.feht_result_fetch.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.feht_result_fetch_device.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.feht_merge_sorted_device.argtypes = [vp, C.c_int, P(i64), P(vp), P(vp), C.c_int, P(vp)]
    L.feht_scatter_rows_device.argtypes = [vp, vp, vp, vp, i64, i32]
    L.feht_host_run.argtypes = [C.c_char_p, i64, C.c_char_p, i64, C.c_char_p, C.c_char_p, C.c_char, C.c_int,
                                C.c_int, f64, C.c_int, C.c_int, P(vp), P(i64), C.c_char_p, C.c_int]
    L.feht_host_plan.argtypes = [C.c_char_p, i64, C.c_char_p, i64, C.c_char_p, C.c_char_p, C.c_char, C.c_int,
                                 P(vp), P(i64), C.c_char_p, C.c_int]
    L.feht_host_show_double.argtypes = [f64, C.c_char_p, C.c_int]
    L.feht_host_free.argtypes = [vp]
    L.feht_host_free.restype = None
    L.feht_set_row_base.argtypes = [vp, i64]
    L.feht_set_comparison_window.argtypes = [vp, i64, i64]
    L.feht_last_timings.argtypes = [vp, P(f64), C.c_int]
    L.feht_merge_sorted.argtypes = [C.c_int, P(i64), P(vp), P(vp), C.c_int, vp, vp]
    L.feht_choose_table.argtypes = [vp, i64]
    L.feht_choose_table.restype = i64
    L.feht_eval_tables.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class FehtHostError(RuntimeError):
    """The reference host calls `error`; the C++ host mirror returns the same message."""


def _host_call(fn, *args) -> bytes:
    out, n = C.c_void_p(), C.c_int64()
    err = C.create_string_buffer(1024)
    rc = fn(*args, C.byref(out), C.byref(n), err, len(err))
    if rc != OK:
        raise FehtHostError(err.value.decode(errors="replace"))
    try:
        return C.string_at(out, n.value)
    finally:
        lib().feht_host_free(out)


def host_run(metadata: bytes, data: bytes, one="all all", two="all all", delimiter=b"\t", mode="binary",
             correction="bonferroni", ratio_filter=0.0, device=0, engine=ENGINE_AUTO) -> bytes:
    """app/Main.hs:13-31 through the C++ host mirror + the CUDA hot path: the bytes feht prints."""
    return _host_call(lib().feht_host_run, metadata, len(metadata), data, len(data), one.encode(), two.encode(),
                      delimiter, int(mode == "snp"), int(correction == "bonferroni"), float(ratio_filter),
                      device, engine)


def host_plan(metadata: bytes, data: bytes, one="all all", two="all all", delimiter=b"\t", mode="binary") -> bytes:
    """Parse + plan only (no GPU)."""
    return _host_call(lib().feht_host_plan, metadata, len(metadata), data, len(data), one.encode(), two.encode(),
                      delimiter, int(mode == "snp"))


def host_show_double(x: float) -> str:
    buf = C.create_string_buffer(64)
    n = lib().feht_host_show_double(float(x), buf, len(buf))
    if n < 0:
        raise FehtHostError("show buffer too small")
    return buf.value.decode()


def pack_chars_host(rows, n_samples: int, snp: bool = False, n_threads: int = 0):
    """feht_pack_chars_host: rows (2-D uint8 or list of bytes) -> (value, valid) or (hi, lo, valid) uint64 planes."""
    flat, off = Context._rows_to_flat(rows)
    n = len(off) - 1
    w = (n_samples + 63) // 64
    planes = [np.zeros((n, w), dtype=np.uint64) for _ in range(3 if snp else 2)]
    p2 = _ptr(planes[2]) if snp else None
    rc = lib().feht_pack_chars_host(_ptr(flat), _ptr(off), n, n_samples, 1 if snp else 0, w, _ptr(planes[0]),
                                    _ptr(planes[1]), p2, n_threads)
    if rc != 0:
        raise FehtError(rc, "feht_pack_chars_host")
    return tuple(planes)


def version() -> str:
    return lib().feht_version().decode()


def choose_table() -> np.ndarray:
    n = lib().feht_choose_table(None, 0)
    out = np.empty(n, dtype=np.float64)
    lib().feht_choose_table(_ptr(out), n)
    return out


def merge_sorted(keys, ranks, sort_key=SORT_ABS_RATIO_DESC):
    """Host k-way merge of sorted shards -> (shard, index) arrays in merged order."""
    n_shards = len(keys)
    keys = [np.ascontiguousarray(k, dtype=np.float64) for k in keys]
    ranks = [np.ascontiguousarray(r, dtype=np.int32) for r in ranks]
    n = (C.c_int64 * max(n_shards, 1))(*[len(k) for k in keys])
    kp = (C.c_void_p * max(n_shards, 1))(*[k.ctypes.data for k in keys])
    rp = (C.c_void_p * max(n_shards, 1))(*[r.ctypes.data for r in ranks])
    total = int(sum(len(k) for k in keys))
    out_s = np.empty(total, dtype=np.int32)
    out_i = np.empty(total, dtype=np.int64)
    rc = lib().feht_merge_sorted(n_shards, n, kp, rp, sort_key, _ptr(out_s), _ptr(out_i))
    if rc != OK:
        raise FehtError(rc, "feht_merge_sorted")
    return out_s, out_i


class Context:
    """One feht_ctx: one GPU (device = an int) or a multi-device handle (device = a list of ids, feht_create_multi)."""

    def __init__(self, device=0):
        self._L = lib()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*[int(d) for d in device])
            rc = self._L.feht_create_multi(C.byref(h), ids, len(device))
        else:
            rc = self._L.feht_create(C.byref(h), device)
        if rc != OK:
            raise FehtError(rc, self._L.feht_last_error(None).decode())
        self._h = h

    @property
    def num_devices(self) -> int:
        return int(self._L.feht_num_devices(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.feht_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc < 0:
            raise FehtError(int(rc), self._L.feht_last_error(self._h).decode())
        return rc

    # ---- data
    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._L.feht_set_stream(self._h, cuda_stream_ptr))

    def set_fet_mode(self, mode: int):
        """FET_TWO_TAIL (default) or FET_ONE_TAIL (FET.hs:82,101-107)."""
        self._check(self._L.feht_set_fet_mode(self._h, mode))

    def set_host_threads(self, n: int):
        """Host threads that transcode part of a chars upload (0 = DMA + GPU pack only, < 0 = default)."""
        self._check(self._L.feht_set_host_threads(self._h, n))

    @property
    def last_upload_host_rows(self) -> int:
        return int(self._L.feht_last_upload_host_rows(self._h))

    def set_samples(self, n: int):
        self._check(self._L.feht_set_samples(self._h, n))

    def reserve_rows(self, n: int):
        self._check(self._L.feht_reserve_rows(self._h, n))

    def clear_rows(self):
        self._check(self._L.feht_clear_rows(self._h))

    @staticmethod
    def _rows_to_flat(rows):
        """list of bytes rows (or a 2-D uint8 array) -> (flat uint8, int64 offsets)."""
        if isinstance(rows, np.ndarray) and rows.ndim == 2:
            flat = np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1)
            off = np.arange(rows.shape[0] + 1, dtype=np.int64) * rows.shape[1]
            return flat, off
        lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
        off = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        flat = np.frombuffer(b"".join(bytes(r) for r in rows), dtype=np.uint8)
        if flat.size == 0:
            flat = np.zeros(1, dtype=np.uint8)
        return np.ascontiguousarray(flat), off

    def add_rows_chars(self, rows, name_rank=None):
        flat, off = self._rows_to_flat(rows)
        nr = _i32(name_rank)
        self._check(self._L.feht_add_rows_chars(self._h, _ptr(flat), _ptr(off), len(off) - 1, _ptr(nr)))

    def add_rows_chars_raw(self, flat_ptr: int, off_ptr: int, n_rows: int, name_rank_ptr: int | None = None):
        """Pointers straight through (bench: pinned torch buffers)."""
        self._check(self._L.feht_add_rows_chars(self._h, flat_ptr, off_ptr, n_rows, name_rank_ptr))

    def add_snp_chars(self, rows, name_rank=None):
        flat, off = self._rows_to_flat(rows)
        nr = _i32(name_rank)
        self._check(self._L.feht_add_snp_chars(self._h, _ptr(flat), _ptr(off), len(off) - 1, _ptr(nr)))

    def add_rows_text(self, text, delimiter: str = "\t", snp: bool = False):
        """Body of the data file (bytes / pinned buffer address + length) -> rows on the device; returns
        (line_begin, name_len) of the lines (feht_add_rows_text + feht_text_lines)."""
        if isinstance(text, tuple):
            ptr, n = text
        else:
            buf = np.frombuffer(text, dtype=np.uint8)
            self._keep_text = buf
            ptr, n = (buf.ctypes.data if len(buf) else None), len(buf)
        nl = C.c_int64(0)
        self._check(self._L.feht_add_rows_text(self._h, ptr, n, delimiter.encode()[0:1], 1 if snp else 0, C.byref(nl)))
        begin = np.empty(nl.value, np.int64)
        name_len = np.empty(nl.value, np.int32)
        self._check(self._L.feht_text_lines(self._h, _ptr(begin), _ptr(name_len), nl.value))
        return begin, name_len

    def set_name_rank(self, name_rank):
        r = _i32(name_rank)
        self._check(self._L.feht_set_name_rank(self._h, _ptr(r), len(r)))

    def add_rows_packed(self, value, valid, name_rank=None):
        value = np.ascontiguousarray(value, dtype=np.uint64)
        valid = np.ascontiguousarray(valid, dtype=np.uint64)
        assert value.shape == valid.shape and value.ndim == 2
        nr = _i32(name_rank)
        self._check(self._L.feht_add_rows_packed(self._h, _ptr(value), _ptr(valid), value.shape[0],
                                                 value.shape[1], _ptr(nr)))

    def add_rows_packed_raw(self, value_ptr: int, valid_ptr: int, n_rows: int, stride_words: int):
        self._check(self._L.feht_add_rows_packed(self._h, value_ptr, valid_ptr, n_rows, stride_words, None))

    def add_snp_packed(self, hi, lo, valid, name_rank=None):
        hi = np.ascontiguousarray(hi, dtype=np.uint64)
        lo = np.ascontiguousarray(lo, dtype=np.uint64)
        valid = np.ascontiguousarray(valid, dtype=np.uint64)
        nr = _i32(name_rank)
        self._check(self._L.feht_add_snp_packed(self._h, _ptr(hi), _ptr(lo), _ptr(valid), hi.shape[0],
                                                hi.shape[1], _ptr(nr)))

    def generate_synthetic(self, spec: SynthSpec, snp_mode: bool, first_row: int, n_rows: int):
        self._check(self._L.feht_generate_synthetic(self._h, C.byref(spec), int(snp_mode), first_row, n_rows))

    def read_plane(self, plane: int, row0: int, n: int) -> np.ndarray:
        w = self.row_stride_words
        out = np.empty((n, w), dtype=np.uint64)
        self._check(self._L.feht_read_plane(self._h, plane, row0, n, _ptr(out)))
        return out

    @property
    def row_stride_words(self) -> int:
        return int(self._L.feht_row_stride_words(self._h))

    @property
    def num_rows(self) -> int:
        return int(self._L.feht_num_rows(self._h))

    # ---- comparisons
    def add_comparison(self, g1_cols, g2_cols) -> int:
        g1, g2 = _i32(g1_cols), _i32(g2_cols)
        return int(self._check(self._L.feht_add_comparison(self._h, _ptr(g1), len(g1), _ptr(g2), len(g2))))

    def add_category(self, value_of_sample, n_values: int) -> int:
        v = _i32(value_of_sample)
        return int(self._check(self._L.feht_add_category(self._h, _ptr(v), n_values)))

    def clear_comparisons(self):
        self._check(self._L.feht_clear_comparisons(self._h))

    @property
    def num_comparisons(self) -> int:
        return int(self._L.feht_num_comparisons(self._h))

    def comparison_info(self, c: int):
        k, cat, v1, v2 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._L.feht_comparison_info(self._h, c, C.byref(k), C.byref(cat), C.byref(v1), C.byref(v2)))
        return k.value, cat.value, v1.value, v2.value

    def set_count_engine(self, engine: int):
        self._check(self._L.feht_set_count_engine(self._h, engine))

    def set_comparison_window(self, first: int = 0, count: int = -1):
        """Emit only comparisons [first, first + count) in the following runs (count < 0: all)."""
        self._check(self._L.feht_set_comparison_window(self._h, first, count))

    def set_row_base(self, base: int):
        self._check(self._L.feht_set_row_base(self._h, base))

    # ---- run
    def run(self, correction=CORRECTION_BONFERRONI, ratio_filter=0.0, sort_key=SORT_ABS_RATIO_DESC,
            bonferroni_n=0):
        self._check(self._L.feht_run(self._h, correction, float(ratio_filter), sort_key, bonferroni_n))

    def run_async(self, correction=CORRECTION_BONFERRONI, ratio_filter=0.0, sort_key=SORT_ABS_RATIO_DESC,
                  bonferroni_n=0):
        """Enqueue the run on the context's stream; result / timing accessors synchronise."""
        self._check(self._L.feht_run_async(self._h, correction, float(ratio_filter), sort_key, bonferroni_n))

    def result_count(self, c: int) -> int:
        return int(self._check(self._L.feht_result_count(self._h, c)))

    def result_total(self) -> int:
        return int(self._check(self._L.feht_result_total(self._h)))

    def fetch(self, c: int) -> dict:
        n = self.result_count(c)
        out = {
            "row": np.empty(n, np.int64), "goa": np.empty(n, np.int32), "gob": np.empty(n, np.int32),
            "gta": np.empty(n, np.int32), "gtb": np.empty(n, np.int32), "p": np.empty(n, np.float64),
            "ratio": np.empty(n, np.float64), "name_rank": np.empty(n, np.int32),
        }
        self._check(self._L.feht_result_fetch(self._h, c, _ptr(out["row"]), _ptr(out["goa"]), _ptr(out["gob"]),
                                              _ptr(out["gta"]), _ptr(out["gtb"]), _ptr(out["p"]),
                                              _ptr(out["ratio"]), _ptr(out["name_rank"])))
        return out

    def fetch_all(self) -> dict:
        """Rows of every comparison, back to back in comparison order (comparison = -1)."""
        n = self.result_total()
        out = {
            "row": np.empty(n, np.int64), "goa": np.empty(n, np.int32), "gob": np.empty(n, np.int32),
            "gta": np.empty(n, np.int32), "gtb": np.empty(n, np.int32), "p": np.empty(n, np.float64),
            "ratio": np.empty(n, np.float64), "name_rank": np.empty(n, np.int32),
        }
        self._check(self._L.feht_result_fetch(self._h, -1, _ptr(out["row"]), _ptr(out["goa"]), _ptr(out["gob"]),
                                              _ptr(out["gta"]), _ptr(out["gtb"]), _ptr(out["p"]),
                                              _ptr(out["ratio"]), _ptr(out["name_rank"])))
        return out

    def result_counts(self) -> np.ndarray:
        return np.array([self.result_count(c) for c in range(self.num_comparisons)], dtype=np.int64)

    def fetch_raw(self, c: int, ptrs: dict):
        """ptrs: name -> host address (pinned buffers owned by the caller)."""
        g = ptrs.get
        self._check(self._L.feht_result_fetch(self._h, c, g("row"), g("goa"), g("gob"), g("gta"), g("gtb"),
                                              g("p"), g("ratio"), g("name_rank")))

    def fetch_device(self, c: int, ptrs: dict):
        """ptrs: name -> DEVICE address on this context's GPU (caller-owned buffers)."""
        g = ptrs.get
        self._check(self._L.feht_result_fetch_device(self._h, c, g("row"), g("goa"), g("gob"), g("gta"),
                                                     g("gtb"), g("p"), g("ratio"), g("name_rank")))

    def merge_sorted_device(self, sizes, key_ptrs, rank_ptrs, dest_ptrs, sort_key=SORT_ABS_RATIO_DESC):
        """K5: device k-way merge; all pointers are device addresses on this context's GPU."""
        k = len(sizes)
        n = (C.c_int64 * max(k, 1))(*[int(x) for x in sizes])
        kp = (C.c_void_p * max(k, 1))(*key_ptrs)
        rp = (C.c_void_p * max(k, 1))(*rank_ptrs)
        dp = (C.c_void_p * max(k, 1))(*dest_ptrs)
        self._check(self._L.feht_merge_sorted_device(self._h, k, n, kp, rp, sort_key, dp))

    def merge_shards_device(self, shards, n_comparisons: int, out: dict, out_seg_off: np.ndarray,
                            sort_key=SORT_ABS_RATIO_DESC):
        """K5b (feht_merge_shards_device).  shards: list of (dict name -> device address, host int64 seg_off array);
        out: dict name -> device address of the merged columns; out_seg_off: host int64 [n_comparisons + 1]."""
        arr = (ShardRows * max(len(shards), 1))()
        keep = []
        for i, (ptrs, seg) in enumerate(shards):
            seg = np.ascontiguousarray(seg, dtype=np.int64)
            keep.append(seg)
            for k in ("row", "goa", "gob", "gta", "gtb", "p", "ratio", "name_rank"):
                setattr(arr[i], k, ptrs.get(k))
            arr[i].seg_off = seg.ctypes.data
        o = ShardRows()
        for k in ("row", "goa", "gob", "gta", "gtb", "p", "ratio", "name_rank"):
            setattr(o, k, out.get(k))
        assert out_seg_off.dtype == np.int64 and len(out_seg_off) == n_comparisons + 1
        o.seg_off = out_seg_off.ctypes.data
        self._check(self._L.feht_merge_shards_device(self._h, len(shards), n_comparisons, arr, sort_key, C.byref(o)))

    def scatter_rows_device(self, src_ptr: int, dst_ptr: int, dest_ptr: int, n: int, elem_bytes: int):
        self._check(self._L.feht_scatter_rows_device(self._h, src_ptr, dst_ptr, dest_ptr, n, elem_bytes))

    def timings(self) -> dict:
        buf = (C.c_double * 9)()
        self._check(self._L.feht_last_timings(self._h, buf, 9))
        keys = ["count_ms", "test_ms", "order_ms", "total_ms", "launches", "count_launches", "count_bytes",
                "engine", "merge_ms"]
        return dict(zip(keys, list(buf)))

    def timings_accumulated(self, reset: bool = False) -> dict:
        buf = (C.c_double * 7)()
        self._check(self._L.feht_timings_accumulated(self._h, buf, 7, 1 if reset else 0))
        keys = ["runs", "count_ms", "test_ms", "order_ms", "total_ms", "launches", "count_launches"]
        return dict(zip(keys, list(buf)))

    def eval_tables(self, goa, gob, gta, gtb):
        a, b, c_, d = _i32(goa), _i32(gob), _i32(gta), _i32(gtb)
        n = len(a)
        p = np.empty(n, np.float64)
        r = np.empty(n, np.float64)
        self._check(self._L.feht_eval_tables(self._h, _ptr(a), _ptr(b), _ptr(c_), _ptr(d), n, _ptr(p), _ptr(r)))
        return p, r
