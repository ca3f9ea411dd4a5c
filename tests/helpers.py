"""Shared helpers for the parity tests (CPU side)."""
from __future__ import annotations

import numpy as np


def make_spec(cls, n_samples, seed=0xFE47, missing=0.01, planted_per_million=1000, split_col=None,
              hi=0.9, lo=0.1, law="beta"):
    """Fill a SynthSpec (oracle's or the library's: same layout).  The per-row frequency table holds
    256 quantiles of Beta(0.3,0.3) (U-shaped pangenome) or of U(0,0.5) (SNP minor)."""
    sp = cls()
    sp.seed = seed
    sp.n_samples = n_samples
    sp.missing_thr = int(missing * 2**32)
    sp.planted_per_million = planted_per_million
    sp.split_col = n_samples // 2 if split_col is None else split_col
    sp.thr_hi = int(hi * 2**32)
    sp.thr_lo = int(lo * 2**32)
    q = (np.arange(256) + 0.5) / 256
    if law == "beta":
        from scipy.stats import beta
        f = beta.ppf(q, 0.3, 0.3)
    else:
        f = 0.5 * q
    thr = np.minimum((f * 2**32).astype(np.uint64), 2**32 - 1)
    for i in range(256):
        sp.freq_thr[i] = int(thr[i])
    return sp


def copy_spec(src, cls):
    import ctypes
    dst = cls()
    ctypes.memmove(ctypes.byref(dst), ctypes.byref(src), ctypes.sizeof(cls))
    return dst


def pack_rows(chars2d: np.ndarray, one=ord("1"), zero=ord("0")):
    """numpy restatement of the packed layout: (value, valid) uint64 [rows][ceil(S/64)]."""
    n, s = chars2d.shape
    w = (s + 63) // 64
    pad = np.zeros((n, w * 64), dtype=np.uint8)
    pad[:, :s] = chars2d
    v = np.packbits((pad == one).astype(np.uint8), axis=1, bitorder="little").view(np.uint64)
    d = np.packbits(((pad == one) | (pad == zero)).astype(np.uint8), axis=1, bitorder="little").view(np.uint64)
    return np.ascontiguousarray(v), np.ascontiguousarray(d)


def pack_snp(chars2d: np.ndarray):
    n, s = chars2d.shape
    w = (s + 63) // 64
    up = chars2d.copy()
    lower = (up >= ord("a")) & (up <= ord("z"))
    up[lower] -= 32
    pad = np.zeros((n, w * 64), dtype=np.uint8)
    pad[:, :s] = up
    isA, isC, isG, isT = (pad == ord(c) for c in "ACGT")
    hi = np.packbits((isG | isT).astype(np.uint8), axis=1, bitorder="little").view(np.uint64)
    lo = np.packbits((isC | isT).astype(np.uint8), axis=1, bitorder="little").view(np.uint64)
    valid = np.packbits((isA | isC | isG | isT).astype(np.uint8), axis=1, bitorder="little").view(np.uint64)
    return hi, lo, valid


def snp_expand(chars2d: np.ndarray):
    """Table.hs:184-205: each site -> rows _a,_t,_c,_g of '1'/'0'/other."""
    n, s = chars2d.shape
    up = chars2d.copy()
    lower = (up >= ord("a")) & (up <= ord("z"))
    up[lower] -= 32
    out = np.empty((n * 4, s), dtype=np.uint8)
    nuc = np.isin(up, np.frombuffer(b"ATCG", dtype=np.uint8))
    for j, ch in enumerate(b"ATCG"):
        row = np.where(up == ch, ord("1"), np.where(nuc, ord("0"), up)).astype(np.uint8)
        out[j::4] = row
    return out


def p_close(got: np.ndarray, want: np.ndarray, chi_branch: np.ndarray):
    """north_star tolerance: FET branch (l < 1000) bit-exact; chi-square branch rel <= 1e-12 or
    abs <= 2^-52 (the reference's p there is quantised to multiples of 2^-53; CUDA's log/exp differ
    from glibc's by <= 1 ulp -- SURVEY.md section 7 'chi-square branch')."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    both_nan = np.isnan(got) & np.isnan(want)
    exact = (got == want) | both_nan
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.abs(got - want) <= 1e-12 * np.abs(want)
    absok = np.abs(got - want) <= 2.0**-52
    ok = np.where(chi_branch, exact | rel | absok, exact)
    return ok


def synth_rows(spec, snp, row0, n_rows, threads=None):
    """oracle.synth_fill for rows (SNP: sites) [row0, row0 + n_rows) on several Python threads (ctypes releases the GIL)."""
    import ctypes
    import os
    from concurrent.futures import ThreadPoolExecutor

    import oracle
    threads = threads or (os.cpu_count() or 1)
    threads = max(1, min(threads, n_rows // 500 or 1))
    if threads == 1:
        return oracle.synth_fill(spec, snp, row0, n_rows)
    out = np.empty((n_rows, spec.n_samples), dtype=np.uint8)
    cuts = [n_rows * t // threads for t in range(threads + 1)]

    def job(t):
        a, b = cuts[t], cuts[t + 1]
        if b > a:
            oracle.lib().fo_synth_fill(ctypes.byref(spec), int(snp), row0 + a, b - a, out[a:b].ctypes.data_as(ctypes.c_void_p))
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(job, range(threads)))
    return out


def check_block_against_oracle(ctx, spec, snp, comps, unit0, n_units, n_total, correction, rf, threads=None):
    """A contiguous block of generated rows (SNP: sites -> 4 allele rows each) through the oracle, compared with the
    rows the GPU returned for the same block: which rows survive the filter, counts and ratio bit-exact, p (exact on the
    FET branch, chi-square tolerance otherwise), and the block's rows in the oracle's relative order.
    comps: {comparison index: (g1, g2)}.  Returns (tests checked, p bit-identical, p compared on the chi-square branch)."""
    import os

    import oracle
    threads = threads or (os.cpu_count() or 1)
    cells = synth_rows(spec, snp, unit0, n_units, threads)
    mult = 4 if snp else 1
    if snp:
        cells = snp_expand(cells)
    lo, hi = unit0 * mult, (unit0 + n_units) * mult
    checked = exact = chi_n = 0
    for ci, (g1, g2) in comps.items():
        want = oracle.run_comparison(cells, g1, g2, correction=correction, bonferroni_n=n_total, n_threads=threads)
        got = ctx.fetch(ci)
        sel = np.flatnonzero((got["row"] >= lo) & (got["row"] < hi))
        keep = np.flatnonzero(rf <= np.abs(want["ratio"]))
        rows = got["row"][sel] - lo
        assert len(rows) == len(keep) and (np.sort(rows) == keep).all(), (ci, len(rows), len(keep))
        for k in ("goa", "gob", "gta", "gtb"):
            assert (got[k][sel] == want[k][rows]).all(), (ci, k)
        assert (got["ratio"][sel] == want["ratio"][rows]).all(), (ci, "ratio")
        l = (want["goa"] + want["gob"] + want["gta"] + want["gtb"])[rows]
        ok = p_close(got["p"][sel], want["p"][rows], l >= 1000)
        assert ok.all(), (ci, got["p"][sel][~ok][:4], want["p"][rows][~ok][:4])
        # relative order of the block's rows in the output = the oracle's order of the block (ties by global row)
        order = oracle.filter_and_order(want["ratio"], None, rf)
        assert (rows == order).all(), (ci, "order inside the block")
        chi = l >= 1000
        both = np.isnan(got["p"][sel]) & np.isnan(want["p"][rows])
        exact += int(((got["p"][sel] == want["p"][rows]) | both)[chi].sum())
        chi_n += int(chi.sum())
        checked += len(want["ratio"])
    return checked, exact, chi_n
