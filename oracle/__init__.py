"""Python handle on the CPU oracle (oracle/feht_oracle.c).  TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference.
Nothing under feht_b200/ imports this package.  Parity status: PINNED (see feht_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "libfeht_oracle.so"
TWO_TAIL, ONE_TAIL = 0, 1


class SynthSpec(C.Structure):
    """fo_synth_spec; field-for-field the same as feht_synth_spec in include/feht_cuda.h."""
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


_lib = None


def build() -> Path:
    subprocess.run(["make", "-C", str(_DIR)], check=True, capture_output=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    src_mtime = max((_DIR / "feht_oracle.c").stat().st_mtime, (_DIR / "feht_oracle.h").stat().st_mtime)
    if not _SO.exists() or _SO.stat().st_mtime < src_mtime:
        build()
    L = C.CDLL(str(_SO))
    d, i, i64, vp = C.c_double, C.c_int, C.c_int64, C.c_void_p
    for name, res, args in [
        ("fo_log1p", d, [d]), ("fo_log_gamma_correction", d, [d]), ("fo_log_beta", d, [d, d]),
        ("fo_log_gamma", d, [d]), ("fo_choose", d, [i, i]), ("fo_incomplete_gamma", d, [d, d]),
        ("fo_hypergeometric_probability", d, [i, i, i, i]),
        ("fo_hypergeometric_cumulative", d, [i, i, i, d]),
        ("fo_chi_squared1_compl_cumulative", d, [d]),
        ("fo_chi_square", d, [i64] * 4), ("fo_fet", d, [i64] * 4 + [i]),
        ("fo_count_char_in_vector_by_indices", i64, [vp, i64, C.c_uint8, vp, i64]),
        ("fo_calculate_ratio", d, [i64] * 4), ("fo_bonferroni", d, [d, i64]),
        ("fo_replace_char", C.c_uint8, [C.c_uint8, C.c_uint8]),
        ("fo_run_comparison", i, [vp, vp, i64, vp, i64, vp, i64, i, i64, vp, vp, vp, vp, vp, vp]),
        ("fo_run_comparison_mt", i, [vp, vp, i64, vp, i64, vp, i64, i, i64, vp, vp, vp, vp, vp, vp, i]),
        ("fo_filter_and_order", i64, [vp, vp, i64, d, vp]),
        ("fo_set_choose_cache", None, [i]),
        ("fo_synth_binary_cell", C.c_uint8, [C.POINTER(SynthSpec), i64, i64]),
        ("fo_synth_snp_cell", C.c_uint8, [C.POINTER(SynthSpec), i64, i64]),
        ("fo_synth_fill", None, [C.POINTER(SynthSpec), i, i64, i64, vp]),
    ]:
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def fet(goa, gob, gta, gtb, mode=TWO_TAIL) -> float:
    return lib().fo_fet(int(goa), int(gob), int(gta), int(gtb), mode)


def choose(n, k) -> float:
    return lib().fo_choose(int(n), int(k))


def calculate_ratio(goa, gob, gta, gtb) -> float:
    return lib().fo_calculate_ratio(int(goa), int(gob), int(gta), int(gtb))


def chi_square(a, b, c, d) -> float:
    return lib().fo_chi_square(int(a), int(b), int(c), int(d))


def chi_squared1_compl_cumulative(x) -> float:
    return lib().fo_chi_squared1_compl_cumulative(float(x))


def bonferroni(p, n) -> float:
    return lib().fo_bonferroni(float(p), int(n))


def count_char_in_vector_by_indices(v: bytes, match: str, indices) -> int:
    buf = np.frombuffer(bytes(v), dtype=np.uint8)
    idx = np.ascontiguousarray(indices, dtype=np.int32)
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
        n = 0
    else:
        n = buf.size
    return int(lib().fo_count_char_in_vector_by_indices(_p(buf), n, ord(match), _p(idx), len(idx)))


def replace_char(match: str, cur: str) -> str:
    return chr(lib().fo_replace_char(ord(match), ord(cur)))


def rows_to_flat(rows):
    if isinstance(rows, np.ndarray) and rows.ndim == 2:
        flat = np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1)
        off = np.arange(rows.shape[0] + 1, dtype=np.int64) * rows.shape[1]
        return flat, off
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
    off = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    flat = np.frombuffer(b"".join(bytes(r) for r in rows), dtype=np.uint8)
    if flat.size == 0:
        flat = np.zeros(1, np.uint8)
    return np.ascontiguousarray(flat), off


def run_comparison(rows, g1, g2, correction=1, bonferroni_n=None, n_threads=1) -> dict:
    """Comparison.hs:86-111 + 238-264 for one comparison over `rows` (list of bytes / 2-D uint8)."""
    flat, off = rows_to_flat(rows)
    n = len(off) - 1
    g1 = np.ascontiguousarray(g1, dtype=np.int32)
    g2 = np.ascontiguousarray(g2, dtype=np.int32)
    out = {k: np.zeros(n, np.int32) for k in ("goa", "gob", "gta", "gtb")}
    out["p"] = np.zeros(n, np.float64)
    out["ratio"] = np.zeros(n, np.float64)
    bn = n if bonferroni_n is None else int(bonferroni_n)
    args = [_p(flat), _p(off), n, _p(g1), len(g1), _p(g2), len(g2), int(correction), bn,
            _p(out["goa"]), _p(out["gob"]), _p(out["gta"]), _p(out["gtb"]), _p(out["p"]), _p(out["ratio"])]
    if n_threads > 1:
        rc = lib().fo_run_comparison_mt(*args, int(n_threads))
    else:
        rc = lib().fo_run_comparison(*args)
    if rc != 0:
        raise IndexError("index out of bounds (V.!)")
    return out


def filter_and_order(ratio, name_rank=None, ratio_filter=0.0) -> np.ndarray:
    ratio = np.ascontiguousarray(ratio, dtype=np.float64)
    nr = None if name_rank is None else np.ascontiguousarray(name_rank, dtype=np.int32)
    out = np.zeros(len(ratio), np.int64)
    n = lib().fo_filter_and_order(_p(ratio), _p(nr), len(ratio), float(ratio_filter), _p(out))
    return out[:n]


def set_choose_cache(enabled: bool):
    lib().fo_set_choose_cache(int(enabled))


def synth_fill(spec: SynthSpec, snp_mode: bool, row0: int, n_rows: int) -> np.ndarray:
    out = np.zeros((n_rows, spec.n_samples), dtype=np.uint8)
    lib().fo_synth_fill(C.byref(spec), int(snp_mode), row0, n_rows, _p(out))
    return out
