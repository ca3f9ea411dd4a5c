"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200.

Bars (BASELINE.json north_star): contingency counts bit-exact; p-values bit-exact on the
hypergeometric branch (l < 1000) and within 1e-12 relative / 2^-52 absolute on the chi-square
branch; ratio bit-exact; row order identical (|ratio| descending, ties by marker-name rank).
"""
import itertools

import numpy as np
import pytest

import oracle
from feht_b200 import capi
from oracle import frontend

from helpers import copy_spec, make_spec, p_close, pack_rows, pack_snp, snp_expand

pytestmark = pytest.mark.gpu
ROOT_DIR = __import__("pathlib").Path(__file__).resolve().parent.parent


def _random_chars(rng, n_rows, n_samples, p_missing=0.05, ragged=False):
    freq = rng.beta(0.3, 0.3, size=(n_rows, 1))
    cells = np.where(rng.random((n_rows, n_samples)) < freq, ord("1"), ord("0")).astype(np.uint8)
    miss = rng.random((n_rows, n_samples)) < p_missing
    cells[miss] = rng.choice(np.frombuffer(b"-?Nx ", dtype=np.uint8), size=int(miss.sum()))
    return cells


def _check_against_oracle(ctx, rows, comps, correction, ratio_filter, name_rank=None, bonf_n=None):
    """rows: 2-D uint8 or list of bytes as held by the reference host; comps: list of (g1, g2)."""
    n_rows = len(rows)
    for ci, pair in enumerate(comps):
        if pair is None:      # not part of this run (comparison window)
            continue
        g1, g2 = pair
        want = oracle.run_comparison(rows, g1, g2, correction=correction,
                                     bonferroni_n=n_rows if bonf_n is None else bonf_n)
        order = oracle.filter_and_order(want["ratio"], name_rank, ratio_filter)
        got = ctx.fetch(ci)
        assert len(got["row"]) == len(order), (ci, len(got["row"]), len(order))
        assert (got["row"] == order).all(), f"comparison {ci}: row order differs"
        for k in ("goa", "gob", "gta", "gtb"):
            assert (got[k] == want[k][order]).all(), (ci, k)
        assert (got["ratio"] == want["ratio"][order]).all(), (ci, "ratio")
        l = (want["goa"] + want["gob"] + want["gta"] + want["gtb"])[order]
        ok = p_close(got["p"], want["p"][order], l >= 1000)
        assert ok.all(), (ci, got["p"][~ok][:5], want["p"][order][~ok][:5])


# ------------------------------------------------------------------------------------ K0 pack
def test_pack_chars_matches_numpy(ctx):
    rng = np.random.default_rng(1)
    for s in (1, 12, 63, 64, 65, 127, 128, 300, 5000):
        cells = _random_chars(rng, 37, s)
        ctx.set_samples(s)
        ctx.add_rows_chars(cells)
        v, d = pack_rows(cells)
        w = v.shape[1]
        gv, gd = ctx.read_plane(0, 0, 37), ctx.read_plane(1, 0, 37)
        assert ctx.row_stride_words % 2 == 0
        assert (gv[:, :w] == v).all() and (gd[:, :w] == d).all()
        assert (gv[:, w:] == 0).all() and (gd[:, w:] == 0).all()


def test_packed_upload_sanitises_padding_and_value_bits(ctx):
    rng = np.random.default_rng(2)
    s = 200
    cells = _random_chars(rng, 19, s)
    v, d = pack_rows(cells)
    # caller stride 6 words with garbage in the tail bits and value bits set where invalid
    vv = np.full((19, 6), np.uint64(0xFFFFFFFFFFFFFFFF))
    dd = np.full((19, 6), np.uint64(0xFFFFFFFFFFFFFFFF))
    vv[:, : v.shape[1]] = v | ~d
    dd[:, : v.shape[1]] = d
    tailmask = np.uint64((1 << (s % 64)) - 1)
    vv[:, -3] |= ~tailmask
    dd[:, 3] |= ~tailmask
    ctx.set_samples(s)
    ctx.add_rows_packed(vv, dd)
    gv, gd = ctx.read_plane(0, 0, 19), ctx.read_plane(1, 0, 19)
    assert (gv[:, :4] == v).all() and (gd[:, :4] == d).all()


def test_device_generator_equals_oracle_generator(ctx):
    for snp in (False, True):
        sp_o = make_spec(oracle.SynthSpec, 333, law="uniform_half" if snp else "beta")
        sp_g = copy_spec(sp_o, capi.SynthSpec)
        ctx.set_samples(333)
        ctx.generate_synthetic(sp_g, snp, 12345, 100)
        cells = oracle.synth_fill(sp_o, snp, 12345, 100)
        if snp:
            hi, lo, valid = pack_snp(cells)
            w = hi.shape[1]
            assert (ctx.read_plane(0, 0, 100)[:, :w] == hi).all()
            assert (ctx.read_plane(1, 0, 100)[:, :w] == valid).all()
            assert (ctx.read_plane(2, 0, 100)[:, :w] == lo).all()
        else:
            v, d = pack_rows(cells)
            w = v.shape[1]
            assert (ctx.read_plane(0, 0, 100)[:, :w] == v).all()
            assert (ctx.read_plane(1, 0, 100)[:, :w] == d).all()


# ------------------------------------------------------------------------ K3 arithmetic alone
def test_fet_exhaustive_small_tables_bit_exact(ctx):
    tabs = [t for t in itertools.product(range(0, 11), repeat=4)]
    a, b, c, d = (np.array(x, dtype=np.int32) for x in zip(*tabs))
    p, r = ctx.eval_tables(a, b, c, d)
    want_p = np.array([oracle.fet(*t) for t in tabs])
    want_r = np.array([oracle.calculate_ratio(*t) for t in tabs])
    assert (p == want_p).all()
    assert (r == want_r).all()


def test_fet_random_hypergeometric_branch_bit_exact(ctx, golden):
    rng = np.random.default_rng(7)
    tabs = []
    for _ in range(1500):
        l = int(rng.integers(50, 1000))
        k = int(rng.integers(1, l))
        if rng.random() < 0.4:
            k = l // 2                      # balanced designs: structural ties
        m = int(rng.integers(0, l + 1))
        lo, hi = max(0, m + k - l), min(m, k)
        goa = int(rng.integers(lo, hi + 1))
        tabs.append((goa, k - goa, m - goa, l - k - m + goa))
    tabs += [(46, 0, 229, 5)]               # exact mathematical tie broken by rounding
    tabs += [tuple(t[:4]) for t in golden["derived_two_tail"]]
    tabs += [(250, 249, 260, 240), (499, 0, 0, 500), (0, 499, 500, 0), (1, 0, 0, 998), (333, 333, 333, 0)]
    a, b, c, d = (np.array(x, dtype=np.int32) for x in zip(*tabs))
    assert ((a + b + c + d) < 1000).all()
    p, r = ctx.eval_tables(a, b, c, d)
    want = np.array([oracle.fet(*t) for t in tabs])
    assert (p == want).all(), np.flatnonzero(p != want)[:10]
    for t, w in zip(golden["derived_two_tail"], p[-5 - len(golden["derived_two_tail"]):-5]):
        assert w == t[4]


def test_fetspec_known_answers_on_device(ctx, golden):
    two = [g for g in golden["fetspec"] if g[4] == "two"]
    a, b, c, d = (np.array([g[i] for g in two], dtype=np.int32) for i in range(4))
    p, _ = ctx.eval_tables(a, b, c, d)
    want = np.array([g[5] for g in two])
    ok = p_close(p, want, (a + b + c + d) >= 1000)
    assert ok.all(), (p, want)
    assert p[0] == want[0] and p[1] == want[1]          # hypergeometric KATs: bit-exact


def test_one_tail_mode_known_answers_and_random_tables(ctx, golden):
    """FETMode OneTail (FET.hs:82,101,105-107): the four OneTail known answers of FETSpec.hs on the device, then
    random tables on both branches against the oracle (hypergeometric cumulative: bit-exact)."""
    ctx.set_fet_mode(capi.FET_ONE_TAIL)
    one = [g for g in golden["fetspec"] if g[4] == "one"]
    assert len(one) == 4
    a, b, c, d = (np.array([g[i] for g in one], dtype=np.int32) for i in range(4))
    p, _ = ctx.eval_tables(a, b, c, d)
    want = np.array([g[5] for g in one])
    assert p_close(p, want, (a + b + c + d) >= 1000).all(), (p, want)
    assert p[0] == want[0] and p[1] == want[1]          # `cumulative d goa`: bit-exact
    rng = np.random.default_rng(17)
    tabs = [t for t in itertools.product(range(0, 6), repeat=4)]
    for _ in range(1200):
        l = int(rng.integers(20, 1000)) if rng.random() < 0.7 else int(rng.integers(1000, 30000))
        k = int(rng.integers(1, l))
        m = int(rng.integers(0, l + 1))
        lo, hi = max(0, m + k - l), min(m, k)
        goa = int(rng.integers(lo, hi + 1))
        tabs.append((goa, k - goa, m - goa, l - k - m + goa))
    a, b, c, d = (np.array(x, dtype=np.int32) for x in zip(*tabs))
    p, _ = ctx.eval_tables(a, b, c, d)
    want = np.array([oracle.fet(*t, mode=oracle.ONE_TAIL) for t in tabs])
    ok = p_close(p, want, (a + b + c + d) >= 1000)
    assert ok.all(), (np.array(tabs)[~ok][:5], p[~ok][:5], want[~ok][:5])
    ctx.set_fet_mode(capi.FET_TWO_TAIL)
    p2, _ = ctx.eval_tables(a[:50], b[:50], c[:50], d[:50])
    assert (p2 == np.array([oracle.fet(*t) for t in tabs[:50]])).all()


def test_chi_square_branch_sweep(ctx):
    rng = np.random.default_rng(9)
    tabs = []
    # around the 999/1000 switch, the Pearson / continued-fraction switch (chi2 = 2) and the
    # zero cut-off (chi2 ~ 72), plus random large tables and zero margins (NaN)
    for l in (1000, 1001, 1500, 4950, 20000, 50000):
        for _ in range(300):
            k = int(rng.integers(1, l))
            m = int(rng.integers(0, l + 1))
            lo, hi = max(0, m + k - l), min(m, k)
            e = m * k / l
            goa = int(np.clip(round(e + rng.normal() * rng.choice([0.3, 1, 3, 9]) * max(1.0, (e * (1 - k / l)) ** 0.5)), lo, hi))
            tabs.append((goa, k - goa, m - goa, l - k - m + goa))
    tabs += [(0, 600, 0, 700), (600, 0, 700, 0), (500, 500, 500, 500), (1000, 0, 0, 1000), (999, 1, 0, 0)]
    a, b, c, d = (np.array(x, dtype=np.int32) for x in zip(*tabs))
    p, r = ctx.eval_tables(a, b, c, d)
    want = np.array([oracle.fet(*t) for t in tabs])
    l = a + b + c + d
    ok = p_close(p, want, l >= 1000)
    assert ok.all(), (np.array(tabs)[~ok][:5], p[~ok][:5], want[~ok][:5])
    assert np.isnan(p[-5]) and np.isnan(p[-4])
    # what was measured, asserted: nearly all p-values are bit-identical; the rest differ through
    # CUDA's log / exp (<= 1 ulp against glibc) and the fast quotient of the two gamma loops, never by more than ONE
    # quantum 2^-53 of `1 - exp g` (the reference's p is a multiple of 2^-53 there), and by <= 1e-13 relative for p > 1e-3
    exact = (p == want) | (np.isnan(p) & np.isnan(want))
    assert float(exact.mean()) >= 0.95, float(exact.mean())
    fin = ~np.isnan(want)
    diff = np.abs(p[fin] - want[fin])
    assert (diff <= 2.0**-53 * (1 + 1e-9)).all() or (diff[diff > 2.0**-53 * (1 + 1e-9)] <= 1e-13 * want[fin][diff > 2.0**-53 * (1 + 1e-9)]).all()
    big = want[fin] > 1e-3
    assert (diff[big] <= 1e-13 * want[fin][big]).all(), float((diff[big] / want[fin][big]).max())


# ------------------------------------------------------------------ K1 + K3 + K4 end to end
@pytest.mark.parametrize("n_samples", [12, 64, 100, 257, 700, 1300, 5000])
def test_random_matrix_explicit_comparisons(ctx, n_samples):
    rng = np.random.default_rng(100 + n_samples)
    n_rows = 257
    cells = _random_chars(rng, n_rows, n_samples)
    cols = rng.permutation(n_samples)
    h = n_samples // 2
    comps = [(cols[:h], cols[h:]),                                     # disjoint, covering
             (cols[: max(1, h // 3)], cols[h: h + max(1, h // 2)]),   # disjoint, partial
             (cols[: max(1, h)], cols[max(0, h // 2): n_samples]),    # overlapping (Comparison.hs:283-284)
             (cols[:1], cols[1:2] if n_samples > 1 else cols[:1])]
    rank = rng.permutation(n_rows).astype(np.int32)
    ctx.set_samples(n_samples)
    ctx.add_rows_chars(cells, rank)
    for g1, g2 in comps:
        ctx.add_comparison(g1, g2)
    for correction, rf in ((1, 0.0), (0, 0.0), (1, 0.3), (0, 1.0)):
        ctx.run(correction, rf)
        _check_against_oracle(ctx, cells, comps, correction, rf, rank)


def test_ragged_rows_and_bounds_error(ctx):
    # rows longer / shorter than S, as Table.hs:170 produces from cells like "1 "
    rows = [b"0101-1010101", b"11110000111100", b"1 01010101010", b"010101010101"]
    ctx.set_samples(12)
    ctx.add_rows_chars(rows)
    g1, g2 = [0, 1, 2, 5], [3, 4, 11]
    ctx.add_comparison(g1, g2)
    ctx.run(capi.CORRECTION_NONE, 0.0)
    _check_against_oracle(ctx, rows, [(g1, g2)], 0, 0.0)
    # a short row + a comparison that indexes past its end = V.! error in the reference
    ctx.add_rows_chars([b"0101"])
    with pytest.raises(capi.FehtError) as e:
        ctx.run(capi.CORRECTION_NONE, 0.0)
    assert e.value.code == capi.E_BOUNDS
    with pytest.raises(IndexError):
        oracle.run_comparison(rows + [b"0101"], g1, g2)


def test_empty_inputs(ctx):
    ctx.set_samples(8)
    ctx.add_comparison([0, 1], [2, 3])
    ctx.run()
    assert ctx.result_count(0) == 0 and ctx.result_total() == 0
    ctx.add_rows_chars([b"01010101"])
    ctx.clear_comparisons()
    ctx.run()
    assert ctx.result_total() == 0
    # a group that is empty for every marker: p = 1, ratio = 0 (FET.hs:99,103; Comparison.hs:74-75)
    ctx.add_comparison([], [0, 1])
    ctx.run(capi.CORRECTION_NONE)
    got = ctx.fetch(0)
    assert got["p"][0] == 1.0 and got["ratio"][0] == 0.0


def test_config1_fixture_end_to_end(ctx, fixture_bytes):
    """BASELINE config 1: data/test_binary.txt + test_metadata.txt, all comparisons, both corrections,
    rows packed from the concatenated cell string exactly as the host produces it."""
    meta = fixture_bytes["test_metadata.txt"]
    for fname, mode in (("test_binary.txt", "binary"), ("test_snps.txt", "snp")):
        pd = frontend.parse_data_file(b"\t", mode, fixture_bytes[fname])
        table = frontend.get_metadata_from_file(b"\t", pd, meta)
        cl = frontend.get_comparison_list(table, frontend.generate_categories("all all", "all all"))
        assert len(cl) == 16
        names = list(pd.gene_vector_map.keys())
        rows = [pd.gene_vector_map[n] for n in names]
        rank = np.argsort(np.argsort(np.array(names, dtype=object))).astype(np.int32)
        ctx.set_samples(12)
        ctx.clear_comparisons()
        ctx.add_rows_chars(rows, rank)
        for comp in cl:
            ctx.add_comparison(comp.g1, comp.g2)
        for correction in (capi.CORRECTION_BONFERRONI, capi.CORRECTION_NONE):
            for rf in (0.0, 0.8, 1.0):
                ctx.run(correction, rf)
                _check_against_oracle(ctx, rows, [(c.g1, c.g2) for c in cl], correction, rf, rank)
                # and the formatted blocks agree with the oracle front-end's
                _, structured = frontend.run_files(meta, fixture_bytes[fname], mode=mode,
                                                   correction="bonferroni" if correction else "none",
                                                   ratio_filter=rf)
                for ci, s in enumerate(structured):
                    got = ctx.fetch(ci)
                    assert [names[i] for i in got["row"]] == s["names"]
                    assert (got["p"] == s["p"]).all()      # l = 12 < 1000: bit-exact


def test_category_api_equals_explicit_comparisons(ctx):
    rng = np.random.default_rng(21)
    n_samples, n_rows = 310, 200
    cells = _random_chars(rng, n_rows, n_samples)
    for n_values in (2, 3, 5, 8):
        vos = rng.integers(-1, n_values, size=n_samples).astype(np.int32)   # -1: no metadata
        ctx.set_samples(n_samples)
        ctx.clear_comparisons()
        ctx.add_rows_chars(cells)
        first = ctx.add_category(vos, n_values)
        assert first == 0
        comps = []
        for i in range(n_values):
            for j in range(i + 1, n_values):
                comps.append((np.flatnonzero(vos == i), np.flatnonzero(vos == j)))
        if n_values > 2:
            for v in range(n_values):
                comps.append((np.flatnonzero(vos == v), np.flatnonzero((vos >= 0) & (vos != v))))
        assert ctx.num_comparisons == len(comps)
        ctx.run(capi.CORRECTION_BONFERRONI, 0.25)
        _check_against_oracle(ctx, cells, comps, 1, 0.25)
        kinds = [ctx.comparison_info(i)[0] for i in range(len(comps))]
        assert kinds.count(2) == (n_values if n_values > 2 else 0)


@pytest.mark.parametrize("n_samples,n_values", [(130, 3), (1000, 7), (2600, 70), (5000, 50), (9100, 2)])
@pytest.mark.parametrize("engine", [capi.ENGINE_GEMM, capi.ENGINE_GEMM_FP4, capi.ENGINE_GEMM_F8])
def test_gemm_engine_counts_equal_oracle(ctx, n_samples, n_values, engine):
    """K2 (tcgen05 GEMM: kind::i8, kind::mxf4, and the one-GEMM e5m2 flavour) against the oracle: same pair / one-vs-rest
    tables, bit-exact, for one and for several passes (70 values = 36 pair-slots > 32 per pass), with an explicit
    comparison mixed in (K1 counts that one), samples that lack the category (-1), and value groups of more than
    4,095 samples (9100 / 2: the e5m2 flavour spreads them over two columns each)."""
    rng = np.random.default_rng(400 + n_values)
    n_rows = 300 if n_values < 50 else 140
    cells = _random_chars(rng, n_rows, n_samples)
    vos = rng.integers(-1 if n_values > 2 else 0, n_values, size=n_samples).astype(np.int32)
    rank = rng.permutation(n_rows).astype(np.int32)
    ctx.set_samples(n_samples)
    ctx.add_rows_chars(cells, rank)
    g1, g2 = np.arange(0, n_samples // 3), np.arange(n_samples // 4, n_samples // 2)
    ctx.add_comparison(g1, g2)
    ctx.add_category(vos, n_values)
    comps = [(g1, g2)]
    for i in range(n_values):
        for j in range(i + 1, n_values):
            comps.append((np.flatnonzero(vos == i), np.flatnonzero(vos == j)))
    for v in range(n_values if n_values > 2 else 0):        # one-vs-rest only for more than two values (Comparison.hs:329)
        comps.append((np.flatnonzero(vos == v), np.flatnonzero((vos >= 0) & (vos != v))))
    assert ctx.num_comparisons == len(comps)
    ctx.set_count_engine(engine)
    ctx.run(capi.CORRECTION_BONFERRONI, 0.4)
    assert ctx.timings()["engine"] == engine
    # every comparison for the small cases, a spread of them for the 2486 / 1276-comparison cases
    pick = range(len(comps)) if len(comps) < 60 else sorted(set(rng.integers(0, len(comps), 60)) | {0, 1, len(comps) - 1})
    for ci in pick:
        g1_, g2_ = comps[ci]
        want = oracle.run_comparison(cells, g1_, g2_, correction=1, bonferroni_n=n_rows)
        order = oracle.filter_and_order(want["ratio"], rank, 0.4)
        got = ctx.fetch(ci)
        assert (got["row"] == order).all(), ci
        for k in ("goa", "gob", "gta", "gtb", "ratio"):
            assert (got[k] == want[k][order]).all(), (ci, k)
        l = (want["goa"] + want["gob"] + want["gta"] + want["gtb"])[order]
        assert p_close(got["p"], want["p"][order], l >= 1000).all(), ci


def test_gemm_and_popcount_engines_agree_on_two_categories(ctx):
    """Two categories in one GEMM pass (C5-style), row count not a multiple of the 256-row tile, rows
    longer than one 1024-sample TMA chunk; dense output (ratio filter 0)."""
    rng = np.random.default_rng(411)
    n_samples, n_rows = 3100, 777
    cells = _random_chars(rng, n_rows, n_samples)
    ctx.set_samples(n_samples)
    ctx.add_rows_chars(cells)
    ctx.add_category(rng.integers(0, 5, size=n_samples).astype(np.int32), 5)
    ctx.add_category(rng.integers(-1, 12, size=n_samples).astype(np.int32), 12)
    res = {}
    for eng in (capi.ENGINE_POPC, capi.ENGINE_GEMM, capi.ENGINE_GEMM_FP4, capi.ENGINE_GEMM_F8, capi.ENGINE_AUTO):
        ctx.set_count_engine(eng)
        ctx.run(capi.CORRECTION_NONE, 0.0)
        res[eng] = [ctx.fetch(ci) for ci in range(ctx.num_comparisons)]
    assert ctx.timings()["engine"] == capi.ENGINE_GEMM_FP4   # AUTO: 3+1+6+1 = 11 pair-slots >= 8 -> the 4-bit GEMM
    for ci in range(ctx.num_comparisons):
        for k in ("row", "goa", "gob", "gta", "gtb", "p", "ratio"):
            a, b, c, d, e8 = (res[e][ci][k] for e in (capi.ENGINE_POPC, capi.ENGINE_GEMM, capi.ENGINE_GEMM_FP4, capi.ENGINE_AUTO, capi.ENGINE_GEMM_F8))
            same = (a == b) | ((a != a) & (b != b))
            assert same.all() and ((b == c) | ((b != b) & (c != c))).all() and ((c == d) | ((c != c) & (d != d))).all(), (ci, k)
            assert ((d == e8) | ((d != d) & (e8 != e8))).all(), (ci, k)


def test_gemm_f8_two_passes_and_fallback(ctx):
    """The e5m2 flavour keeps a category inside one pass of 128 columns: two 70-value categories make two passes;
    a 130-value category does not fit (asking for the flavour is an error; AUTO uses the 4-bit GEMM anyway)."""
    rng = np.random.default_rng(412)
    n_samples, n_rows = 2600, 333
    cells = _random_chars(rng, n_rows, n_samples)
    ctx.set_samples(n_samples)
    ctx.add_rows_chars(cells)
    ctx.add_category(rng.integers(0, 70, size=n_samples).astype(np.int32), 70)
    ctx.add_category(rng.integers(-1, 70, size=n_samples).astype(np.int32), 70)
    res = {}
    for eng in (capi.ENGINE_GEMM_FP4, capi.ENGINE_GEMM_F8):
        ctx.set_count_engine(eng)
        ctx.run(capi.CORRECTION_NONE, 0.3)
        assert ctx.timings()["engine"] == eng
        res[eng] = [ctx.fetch(ci) for ci in range(0, ctx.num_comparisons, 37)]
    for a, b in zip(res[capi.ENGINE_GEMM_FP4], res[capi.ENGINE_GEMM_F8]):
        for k in ("row", "goa", "gob", "gta", "gtb", "ratio"):
            assert (a[k] == b[k]).all(), k
    ctx.set_samples(n_samples)
    ctx.add_rows_chars(cells)
    ctx.clear_comparisons()
    ctx.add_category(rng.integers(0, 130, size=n_samples).astype(np.int32), 130)
    ctx.set_count_engine(capi.ENGINE_AUTO)
    ctx.run(capi.CORRECTION_NONE, 0.5)
    assert ctx.timings()["engine"] == capi.ENGINE_GEMM_FP4
    ctx.set_count_engine(capi.ENGINE_GEMM_F8)
    with pytest.raises(capi.FehtError):
        ctx.run(capi.CORRECTION_NONE, 0.5)
    ctx.set_count_engine(capi.ENGINE_AUTO)


def test_gemm_engine_rejects_snp_and_explicit_only(ctx):
    ctx.set_samples(64)
    ctx.add_rows_chars([b"01" * 32])
    ctx.add_comparison([0, 1], [2, 3])
    ctx.set_count_engine(capi.ENGINE_GEMM)
    with pytest.raises(capi.FehtError):
        ctx.run()
    ctx.set_count_engine(capi.ENGINE_AUTO)
    ctx.run()


def test_snp_packed_equals_expanded_binary_rows(ctx):
    rng = np.random.default_rng(33)
    n_sites, n_samples = 150, 421
    alphabet = np.frombuffer(b"ACGTacgt-N", dtype=np.uint8)
    sites = rng.choice(alphabet, size=(n_sites, n_samples), p=[.3, .2, .2, .2, .02, .02, .02, .02, .01, .01])
    expanded = snp_expand(sites)                       # what the Haskell host would hand over
    vos = rng.integers(0, 4, size=n_samples).astype(np.int32)
    comps = [(np.flatnonzero(vos == i), np.flatnonzero(vos == j)) for i in range(4) for j in range(i + 1, 4)]
    comps += [(np.flatnonzero(vos == v), np.flatnonzero(vos != v)) for v in range(4)]
    # (1) packed planes + category, (2) site characters + category, (3) expanded chars + explicit
    hi, lo, valid = pack_snp(sites)
    for variant in range(3):
        ctx.set_samples(n_samples)
        ctx.clear_comparisons()
        if variant == 0:
            ctx.add_snp_packed(hi, lo, valid)
        elif variant == 1:
            ctx.add_snp_chars(sites)
        else:
            ctx.add_rows_chars(expanded)
        if variant < 2:
            ctx.add_category(vos, 4)
        else:
            for g1, g2 in comps:
                ctx.add_comparison(g1, g2)
        assert ctx.num_rows == 4 * n_sites
        ctx.run(capi.CORRECTION_BONFERRONI, 0.1)
        _check_against_oracle(ctx, expanded, comps, 1, 0.1)


@pytest.mark.parametrize("n_samples,n_sites,n_values", [
    (1, 700, 2), (64, 5000, 3), (129, 3001, 4), (1000, 2500, 5), (5000, 1203, 4), (10000, 641, 4),
    (20000, 300, 3), (40000, 97, 4), (70000, 40, 2)])
def test_snp_count_kernels_shape_sweep(ctx, n_samples, n_sites, n_values):
    """K1 for SNP planes over the shapes that select its variants (count_popc.cu): the TMA-fed ring with 16 or 8 warps,
    several sites per lane group, tiles that end inside the matrix, an odd number of pair-slots, rows too wide for
    the ring (register-staged kernel).  hi / lo planes deliberately carry bits where the call is missing.  Every
    2x2 table of every comparison against numpy popcounts (bit-exact), for a whole category."""
    rng = np.random.default_rng(n_samples * 7 + n_values)
    W = (n_samples + 63) // 64
    hi = rng.integers(0, 2**63, size=(n_sites, W), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n_sites, W), dtype=np.uint64)
    lo = rng.integers(0, 2**63, size=(n_sites, W), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n_sites, W), dtype=np.uint64)
    missing = rng.integers(0, 2**63, size=(n_sites, W), dtype=np.uint64) & rng.integers(0, 2**63, size=(n_sites, W), dtype=np.uint64) \
        & rng.integers(0, 2**63, size=(n_sites, W), dtype=np.uint64)
    valid = ~missing
    if n_samples % 64:
        valid[:, -1] &= np.uint64((1 << (n_samples % 64)) - 1)
    vos = rng.integers(-1, n_values, size=n_samples).astype(np.int32)    # -1: genome without the category
    ctx.set_samples(n_samples)
    ctx.clear_comparisons()
    ctx.add_snp_packed(hi, lo, valid)
    ctx.add_category(vos, n_values)
    ctx.run(capi.CORRECTION_NONE, 0.0)
    # numpy: bits of the four allele rows of every site
    def bits(a):
        return np.unpackbits(a.view(np.uint8), axis=1, bitorder="little")[:, :n_samples].astype(bool)
    bh, bl, bv = bits(hi), bits(lo), bits(valid)
    allele = {0: ~bh & ~bl & bv, 1: bh & bl & bv, 2: ~bh & bl & bv, 3: bh & ~bl & bv}   # _a _t _c _g (A=0 C=1 G=2 T=3)
    P = np.stack([allele[k] for k in range(4)], axis=1).reshape(4 * n_sites, n_samples)
    V = np.repeat(bv, 4, axis=0)
    onehot = np.stack([vos == v for v in range(n_values)], axis=1).astype(np.int64)       # [S][V]
    pos = P.astype(np.int64) @ onehot
    tot = V.astype(np.int64) @ onehot
    assert ctx.num_comparisons == n_values * (n_values - 1) // 2 + (n_values if n_values > 2 else 0)
    for c in range(ctx.num_comparisons):
        kind, _cat, v1, v2 = ctx.comparison_info(c)
        got = ctx.fetch(c)
        order = np.argsort(got["row"])
        assert (got["row"][order] == np.arange(4 * n_sites)).all()
        a, n1 = pos[:, v1], tot[:, v1]
        if kind == 1:    # pair
            b, n2 = pos[:, v2], tot[:, v2]
        else:            # one-vs-rest
            b, n2 = pos.sum(axis=1) - a, tot.sum(axis=1) - n1
        assert (got["goa"][order] == a).all() and (got["gob"][order] == n1 - a).all(), (c, kind, v1, v2)
        assert (got["gta"][order] == b).all() and (got["gtb"][order] == n2 - b).all(), (c, kind, v1, v2)


def _random_table_text(rng, n_lines, n_cols, delim, snp, crlf=True):
    """Body of a data file the way Table.hs reads it, with the quirks that matter: cells wider than one character
    (they shift the later columns), empty cells, '\\r' line ends and stray '\\r', ragged lines, a last line without
    a newline, names of any length (also empty).  Returns (text, names, rows) with rows = what parseDataFile hands on."""
    cells_bin = [b"1", b"0", b"1", b"0", b"-", b"1 ", b"NA", b"", b"?"]
    cells_snp = [b"A", b"C", b"G", b"T", b"a", b"c", b"g", b"t", b"-", b"N", b"AC", b""]
    pool = cells_snp if snp else cells_bin
    weights = np.array([6, 6, 6, 6, 1, 0.3, 0.3, 0.3, 0.3][: len(pool)] + [0.3] * max(0, len(pool) - 9), dtype=float)
    if snp:
        weights = np.array([5, 5, 5, 5, 1, 1, 1, 1, 1, 0.5, 0.3, 0.3])
    weights /= weights.sum()
    d = delim.encode()
    lines, names, rows = [], [], []
    for i in range(n_lines):
        k = n_cols if rng.random() < 0.9 else int(rng.integers(0, n_cols + 3))
        name = (b"m%d" % i) * int(rng.integers(0, 3)) + (b"x" if rng.random() < 0.5 else b"")
        cells = [pool[j] for j in rng.choice(len(pool), size=k, p=weights)]
        line = d.join([name] + cells)
        if line == b"":
            line = b"q"                      # (an empty line is an error: tested separately)
        if crlf and rng.random() < 0.3:
            line = line[: len(line) // 2] + b"\r" + line[len(line) // 2:]
        lines.append(line + (b"\r" if crlf and i % 3 == 0 else b""))
        clean = lines[-1].replace(b"\r", b"").split(d)
        names.append(clean[0])
        rows.append(b"".join(clean[1:]))
    return b"\n".join(lines) + (b"\n" if n_lines % 2 else b""), names, rows


@pytest.mark.parametrize("snp", [False, True])
@pytest.mark.parametrize("delim", ["\t", ","])
def test_text_upload_equals_host_parsed_rows(ctx, snp, delim):
    """feht_add_rows_text (text_load.cu: lines, first cell = name, the rest concatenated, packed on the device) against
    the rows a host parser following Table.hs:267-279,160-191 would hand to feht_add_rows_chars: same planes, same
    line offsets and name lengths, same shortest row (bounds check of V.!)."""
    rng = np.random.default_rng(5 + snp)
    for n_lines, n_cols in ((1, 1), (37, 12), (500, 200), (1200, 1031)):
        text, names, rows = _random_table_text(rng, n_lines, n_cols, delim, snp)
        S = n_cols
        ctx.set_samples(S)
        begin, name_len = ctx.add_rows_text(text, delim, snp)
        assert len(begin) == n_lines
        got = [ctx.read_plane(k, 0, n_lines) for k in range(3 if snp else 2)]
        raw_lines = text.split(b"\n")
        pos = 0
        for i in range(n_lines):
            assert begin[i] == pos, i
            assert name_len[i] == len(raw_lines[i].split(delim.encode())[0]), i
            pos += len(raw_lines[i]) + 1
        ctx.set_samples(S)
        (ctx.add_snp_chars if snp else ctx.add_rows_chars)(rows)
        want = [ctx.read_plane(k, 0, n_lines) for k in range(3 if snp else 2)]
        for g, w in zip(got, want):
            assert (g == w).all()


def test_text_upload_randomised_shapes(ctx):
    """60 random files (1..3000 lines, 1..700 columns, either delimiter, binary and SNP, with and without CR):
    planes of the device text path against the character rows."""
    rng = np.random.default_rng(2024)
    for it in range(60):
        snp = bool(rng.integers(0, 2))
        delim = "\t" if rng.integers(0, 2) else ","
        n_lines = int(rng.choice([1, 2, 31, 32, 33, 257, 1000, 3000]))
        n_cols = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 127, 129, 300, 700]))
        text, _names, rows = _random_table_text(rng, n_lines, n_cols, delim, snp, crlf=bool(rng.integers(0, 2)))
        ctx.set_samples(n_cols)
        begin, name_len = ctx.add_rows_text(text, delim, snp)
        assert len(begin) == n_lines
        got = [ctx.read_plane(k, 0, n_lines) for k in range(3 if snp else 2)]
        ctx.set_samples(n_cols)
        (ctx.add_snp_chars if snp else ctx.add_rows_chars)(rows)
        want = [ctx.read_plane(k, 0, n_lines) for k in range(3 if snp else 2)]
        for k, (g, w) in enumerate(zip(got, want)):
            assert (g == w).all(), (it, snp, delim, n_lines, n_cols, k)


def test_text_upload_errors_and_name_ranks(ctx):
    S = 6
    ctx.set_samples(S)
    with pytest.raises(capi.FehtError) as e:
        ctx.add_rows_text(b"a\t1\t0\n\nb\t1\t1\n", "\t", False)
    assert "Irrefutable pattern" in str(e.value) and ctx.num_rows == 0
    ctx.set_samples(S)
    with pytest.raises(capi.FehtError) as e:
        ctx.add_rows_text(b"s1\tA\tC\t1\n", "\t", True)
    assert "literal" in str(e.value)
    # ranks set afterwards order the ties: three identical rows, ranks 2,0,1
    ctx.set_samples(S)
    ctx.add_rows_text(b"z\t1\t1\t1\t0\t0\t0\ny\t1\t1\t1\t0\t0\t0\nx\t1\t1\t1\t0\t0\t0", "\t", False)
    assert ctx.num_rows == 3
    ctx.set_name_rank(np.array([2, 0, 1], dtype=np.int32))
    ctx.clear_comparisons()
    ctx.add_comparison(np.arange(0, 3), np.arange(3, 6))
    ctx.run(capi.CORRECTION_NONE, 0.0)
    got = ctx.fetch(0)
    assert list(got["row"]) == [1, 2, 0] and list(got["goa"]) == [3, 3, 3] and list(got["gtb"]) == [3, 3, 3]
    # a row shorter than a group's columns is the reference's V.! error
    ctx.set_samples(S)
    ctx.add_rows_text(b"a\t1\t1\t1\t0\t0\t0\nb\t1\t1\n", "\t", False)
    ctx.clear_comparisons()
    ctx.add_comparison(np.arange(0, 3), np.arange(3, 6))
    with pytest.raises(capi.FehtError):
        ctx.run(capi.CORRECTION_NONE, 0.0)


def test_text_upload_many_pieces(cuda_lib):
    """150 MB of text = three pieces of whole lines: planes equal those of the character rows, line offsets
    continue across the pieces."""
    rng = np.random.default_rng(8)
    S = 300
    block_text, _names, block_rows = _random_table_text(rng, 4000, S, "\t", False, crlf=False)
    if not block_text.endswith(b"\n"):
        block_text += b"\n"
    reps = (150 << 20) // len(block_text) + 1
    text = block_text * reps
    n_lines = 4000 * reps
    with capi.Context(0) as c:
        c.set_samples(S)
        begin, name_len = c.add_rows_text(text, "\t", False)
        assert len(begin) == n_lines and begin[4000] == len(block_text) and begin[-4000] == len(block_text) * (reps - 1)
        assert (np.diff(begin) > 0).all()
        got = [c.read_plane(k, n_lines - 4000, 4000) for k in range(2)]
        first = [c.read_plane(k, 0, 4000) for k in range(2)]
        c.set_samples(S)
        c.add_rows_chars(block_rows)
        want = [c.read_plane(k, 0, 4000) for k in range(2)]
    for g, f, w in zip(got, first, want):
        assert (g == w).all() and (f == w).all()


@pytest.mark.parametrize("n_samples,n_rows,n_values", [
    (64, 20000, 4), (129, 9001, 8), (1000, 5000, 5), (5000, 3001, 14), (10000, 500, 9), (40000, 100, 4),
    (70000, 40, 3), (300, 400000, 7)])
def test_popcount_engine_category_shape_sweep(ctx, n_samples, n_rows, n_values):
    """K1 for binary planes with several pair-slots per read of the planes (count_popc.cu: the TMA-fed ring with 4 / 2
    pair-slots per pass, 16 or 8 warps, stages refilled many times over on the long matrix; one pair-slot and rows
    too wide for the ring through the register-staged kernel).  Every 2x2 table of every comparison of a whole
    category against numpy popcounts (bit-exact)."""
    rng = np.random.default_rng(n_samples * 3 + n_values)
    W = (n_samples + 63) // 64
    def words():
        return rng.integers(0, 2**63, size=(n_rows, W), dtype=np.uint64) * np.uint64(2) + \
            rng.integers(0, 2, size=(n_rows, W), dtype=np.uint64)
    value = words()
    valid = words() | words() | words()
    if n_samples % 64:
        valid[:, -1] &= np.uint64((1 << (n_samples % 64)) - 1)
    vos = rng.integers(-1, n_values, size=n_samples).astype(np.int32)
    ctx.set_samples(n_samples)
    ctx.clear_comparisons()
    ctx.add_rows_packed(value, valid)        # value bits outside valid are dropped by the upload
    ctx.add_category(vos, n_values)
    ctx.set_count_engine(capi.ENGINE_POPC)
    ctx.run(capi.CORRECTION_NONE, 0.0)
    ctx.set_count_engine(capi.ENGINE_AUTO)
    def bits(a):
        return np.unpackbits(a.view(np.uint8), axis=1, bitorder="little")[:, :n_samples]
    bv = bits(valid)
    bp = bits(value) & bv
    onehot = np.stack([vos == v for v in range(n_values)], axis=1).astype(np.float32)
    pos = (bp.astype(np.float32) @ onehot).astype(np.int64)      # exact: counts < 2^24
    tot = (bv.astype(np.float32) @ onehot).astype(np.int64)
    for c in sorted(set(rng.integers(0, ctx.num_comparisons, 12)) | {0, ctx.num_comparisons - 1}):
        kind, _cat, v1, v2 = ctx.comparison_info(int(c))
        got = ctx.fetch(int(c))
        order = np.argsort(got["row"])
        assert (got["row"][order] == np.arange(n_rows)).all()
        a, n1 = pos[:, v1], tot[:, v1]
        if kind == 1:
            b, n2 = pos[:, v2], tot[:, v2]
        else:
            b, n2 = pos.sum(axis=1) - a, tot.sum(axis=1) - n1
        assert (got["goa"][order] == a).all() and (got["gob"][order] == n1 - a).all(), (c, kind, v1, v2)
        assert (got["gta"][order] == b).all() and (got["gtb"][order] == n2 - b).all(), (c, kind, v1, v2)


def test_row_sharding_and_merge_equals_single_context(cuda_lib):
    """shard marker rows, pass the global Bonferroni n, k-way merge the sorted shards."""
    rng = np.random.default_rng(55)
    n_samples, n_rows = 500, 1000
    cells = _random_chars(rng, n_rows, n_samples)
    rank = rng.permutation(n_rows).astype(np.int32)
    g1, g2 = np.arange(0, 200), np.arange(250, 500)
    want = oracle.run_comparison(cells, g1, g2, correction=1, bonferroni_n=n_rows)
    order = oracle.filter_and_order(want["ratio"], rank, 0.2)
    shards = []
    bounds = [0, 300, 301, 777, 1000]
    for lo_, hi_ in zip(bounds[:-1], bounds[1:]):
        c = capi.Context(0)
        c.set_samples(n_samples)
        c.set_row_base(lo_)
        c.add_rows_chars(cells[lo_:hi_], rank[lo_:hi_])
        c.add_comparison(g1, g2)
        c.run(1, 0.2, capi.SORT_ABS_RATIO_DESC, n_rows)
        shards.append(c.fetch(0))
        c.close()
    out_s, out_i = capi.merge_sorted([s["ratio"] for s in shards], [s["name_rank"] for s in shards])
    merged_rows = np.array([shards[s]["row"][i] for s, i in zip(out_s, out_i)])
    merged_p = np.array([shards[s]["p"][i] for s, i in zip(out_s, out_i)])
    assert (merged_rows == order).all()
    assert (merged_p == want["p"][order]).all()


def test_device_kway_merge_equals_host_merge(cuda_lib):
    """K5 (feht_merge_sorted_device + feht_scatter_rows_device): shards fetched into device buffers,
    merged on the device, equal to the host k-way merge and to the oracle order; ties in |ratio|
    across shards are broken by name rank."""
    import torch
    rng = np.random.default_rng(56)
    n_samples, n_rows = 120, 3000                     # few samples -> many tied ratios
    cells = _random_chars(rng, n_rows, n_samples)
    rank = rng.permutation(n_rows).astype(np.int32)
    g1, g2 = np.arange(0, 50), np.arange(60, 120)
    want = oracle.run_comparison(cells, g1, g2, correction=1, bonferroni_n=n_rows)
    order = oracle.filter_and_order(want["ratio"], rank, 0.05)
    bounds = [0, 700, 701, 1800, 3000]
    dev = torch.device("cuda", 0)
    shards, sizes, ctxs = [], [], []
    for lo_, hi_ in zip(bounds[:-1], bounds[1:]):
        c = capi.Context(0)
        c.set_samples(n_samples)
        c.set_row_base(lo_)
        c.add_rows_chars(cells[lo_:hi_], rank[lo_:hi_])
        c.add_comparison(g1, g2)
        c.run(1, 0.05, capi.SORT_ABS_RATIO_DESC, n_rows)
        n = c.result_count(0)
        t = {"row": torch.empty(max(n, 1), dtype=torch.int64, device=dev),
             "ratio": torch.empty(max(n, 1), dtype=torch.float64, device=dev),
             "p": torch.empty(max(n, 1), dtype=torch.float64, device=dev),
             "name_rank": torch.empty(max(n, 1), dtype=torch.int32, device=dev)}
        c.fetch_device(0, {k: v.data_ptr() for k, v in t.items()})
        shards.append(t); sizes.append(n); ctxs.append(c)
    total = sum(sizes)
    assert total == len(order)
    dest = [torch.empty(max(n, 1), dtype=torch.int64, device=dev) for n in sizes]
    c0 = ctxs[0]
    c0.merge_sorted_device(sizes, [t["ratio"].data_ptr() for t in shards],
                           [t["name_rank"].data_ptr() for t in shards], [d.data_ptr() for d in dest])
    merged_row = torch.full((total,), -1, dtype=torch.int64, device=dev)
    merged_p = torch.zeros(total, dtype=torch.float64, device=dev)
    for t, d, n in zip(shards, dest, sizes):
        c0.scatter_rows_device(t["row"].data_ptr(), merged_row.data_ptr(), d.data_ptr(), n, 8)
        c0.scatter_rows_device(t["p"].data_ptr(), merged_p.data_ptr(), d.data_ptr(), n, 8)
    torch.cuda.synchronize()
    assert (merged_row.cpu().numpy() == order).all()
    assert (merged_p.cpu().numpy() == want["p"][order]).all()
    # and the host merge gives the same (shard, index) sequence
    hs = [{k: v[:n].cpu().numpy() for k, v in t.items()} for t, n in zip(shards, sizes)]
    out_s, out_i = capi.merge_sorted([h["ratio"] for h in hs], [h["name_rank"] for h in hs])
    host_rows = np.array([hs[s_]["row"][i_] for s_, i_ in zip(out_s, out_i)])
    assert (host_rows == order).all()
    for c in ctxs:
        c.close()


def test_pvalue_sort_key(ctx):
    rng = np.random.default_rng(77)
    cells = _random_chars(rng, 300, 90)
    ctx.set_samples(90)
    ctx.add_rows_chars(cells)
    ctx.add_comparison(np.arange(0, 40), np.arange(40, 90))
    ctx.run(capi.CORRECTION_NONE, 0.0, capi.SORT_PVALUE_ASC)
    got = ctx.fetch(0)
    keys = list(zip(got["p"], got["row"]))
    assert keys == sorted(keys)


# ------------------------------------------------------------- BASELINE sizes: properties
def test_config2_full_size_properties(ctx):
    """Config 2 (1M markers x 5k genomes, one comparison): too big for the oracle, so check
    size-independent properties plus an oracle comparison on rows regenerated from the counter-based
    generator."""
    S, R = 5000, 1_000_000
    sp_o = make_spec(oracle.SynthSpec, S)
    sp_g = copy_spec(sp_o, capi.SynthSpec)
    ctx.set_samples(S)
    ctx.generate_synthetic(sp_g, False, 0, R)
    g1, g2 = np.arange(0, 2500), np.arange(2500, 5000)
    ctx.add_comparison(g1, g2)
    ctx.run(capi.CORRECTION_BONFERRONI, 0.0)
    got = ctx.fetch(0)
    assert len(got["row"]) == R
    assert (np.sort(got["row"]) == np.arange(R)).all()                     # a permutation of the rows
    ar = np.abs(got["ratio"])
    assert (ar[:-1] >= ar[1:]).all()                                        # |ratio| descending
    ties = ar[:-1] == ar[1:]
    assert (got["row"][:-1][ties] < got["row"][1:][ties]).all()             # ties by rank (= row)
    n1 = got["goa"] + got["gob"]
    n2 = got["gta"] + got["gtb"]
    assert n1.max() <= 2500 and n2.max() <= 2500 and (n1 + n2).min() > 4500  # ~1% missing
    with np.errstate(invalid="ignore", divide="ignore"):
        r = got["goa"] / n1 - got["gta"] / n2
    assert (r == got["ratio"]).all()                                        # ratio is two IEEE divisions
    # l >= 1000 with an empty column margin: chiSquare = 0/0 = NaN, and NaN survives Bonferroni
    # (FET.hs:132-133; Comparison.hs:262 `NaN > 1` is False) -- SURVEY.md Appendix A.4
    zero_margin = ((got["goa"] + got["gta"]) == 0) | ((got["gob"] + got["gtb"]) == 0)
    assert (np.isnan(got["p"]) == zero_margin).all()
    assert ((got["p"] >= 0) & (got["p"] <= 1))[~zero_margin].all()
    # planted rows (f = 0.9 vs 0.1) come out on top
    assert ar[0] > 0.7
    # oracle on a sample of rows, regenerated on the CPU
    sample = np.concatenate([got["row"][:40], got["row"][-20:], np.array([0, 1, R - 1, 123456])])
    for row in sample:
        cells = oracle.synth_fill(sp_o, False, int(row), 1)
        want = oracle.run_comparison(cells, g1, g2, correction=1, bonferroni_n=R)
        i = int(np.flatnonzero(got["row"] == row)[0])
        assert (got["goa"][i], got["gob"][i], got["gta"][i], got["gtb"][i]) == \
            (want["goa"][0], want["gob"][0], want["gta"][0], want["gtb"][0])
        assert got["ratio"][i] == want["ratio"][0]
        assert p_close(got["p"][i:i + 1], want["p"], np.array([True])).all()
    t = ctx.timings()
    assert t["count_launches"] >= 1 and t["total_ms"] > 0


# ------------------------------------------------------------- the C++ host mirror, end to end
@pytest.mark.parametrize("data_name,meta_name,delim,mode", [
    ("test_binary.txt", "test_metadata.txt", b"\t", "binary"),
    ("test_snps.txt", "test_metadata.txt", b"\t", "snp"),
    ("test_snps.csv", "test_metadata.csv", b",", "snp"),
])
def test_host_run_prints_what_feht_prints(cuda_lib, fixture_bytes, data_name, meta_name, delim, mode):
    """app/Main.hs:13-31 through feht_host_run: byte-identical to the oracle front-end's blocks (which
    reproduce README.md:116-139,170-180,214-222), for every option combination of the CLI."""
    data, meta = fixture_bytes[data_name], fixture_bytes[meta_name]
    for one, two in (("all all", "all all"), ("group A", "group B"), ("position sideways", "all"),
                     ("group A B", "position up")):
        for correction in ("bonferroni", "none"):
            for rf in (0.0, 0.5, 0.8, 1.0):
                blocks, _ = frontend.run_files(meta, data, one=one, two=two, delimiter=delim, mode=mode,
                                               correction=correction, ratio_filter=rf)
                want = b"".join(b + b"\n" for b in blocks) + b"Done\n"
                for engine in (capi.ENGINE_POPC, capi.ENGINE_AUTO):
                    got = capi.host_run(meta, data, one, two, delim, mode, correction, rf, engine=engine)
                    assert got == want, (one, two, correction, rf, engine)


def test_host_run_device_text_path_equals_host_parser(cuda_lib, fixture_bytes, monkeypatch):
    """feht_host_run sends the body of the data file to the device as text (feht_add_rows_text) and falls back to
    parse_data_file when marker names repeat or SNP text holds literal 0/1; FEHT_HOST_TEXT=0 forces the host
    parser.  Same bytes out either way: the fixtures, CRLF line ends, a cell wider than one character, a ragged
    line, a repeated name, a literal '1' in SNP data."""
    meta = fixture_bytes["test_metadata.txt"]
    binary = fixture_bytes["test_binary.txt"]
    snps = fixture_bytes["test_snps.txt"]
    blines = binary.split(b"\n")
    variants = [
        ("binary", binary),
        ("binary", binary.replace(b"\n", b"\r\n")),
        ("binary", b"\n".join(blines[:3] + [blines[3] + b"\t1"] + blines[4:])),            # one line longer
        ("binary", b"\n".join(blines[:5] + [blines[5].rsplit(b"\t", 2)[0]] + blines[6:])),  # one line shorter
        ("binary", b"\n".join(blines[:8] + [blines[2]] + blines[8:])),                       # a name twice
        ("snp", snps),
        ("snp", snps.replace(b"\n", b"\r\n")),
        ("snp", snps.replace(b"\tA\t", b"\t1\t", 1)),                                      # literal 1
    ]
    for mode, data in variants:
        for one, two, rf in (("all all", "all all", 0.0), ("group A", "group B", 0.5)):
            outs = []
            for env in ("1", "0"):
                monkeypatch.setenv("FEHT_HOST_TEXT", env)
                try:
                    outs.append(capi.host_run(meta, data, one, two, b"\t", mode, "none", rf))
                except capi.FehtHostError as e:
                    outs.append(("error", str(e)))
            assert outs[0] == outs[1], (mode, one, two, rf, outs[0][:200], outs[1][:200])
    monkeypatch.delenv("FEHT_HOST_TEXT")


def test_host_run_readme_rows(cuda_lib, fixture_bytes):
    """README.md:133: `binary9 2 0 1 4 ... 0.8` in the position sideways-vs-up block."""
    out = capi.host_run(fixture_bytes["test_metadata.txt"], fixture_bytes["test_binary.txt"],
                        "position sideways", "position up", ratio_filter=0.8)
    assert b"Group1 category: position Group1: sideways\nGroup2 category: position Group2: up\n---\n" in out
    assert b"binary9\t2\t0\t1\t4\t1.0\t0.8\n" in out
    assert out.endswith(b"-#]\n\nDone\n")


def test_cli_binary(cuda_lib, tmp_path):
    import subprocess
    from conftest import DATA, ROOT
    exe = ROOT / "feht_b200" / "bin" / "feht-b200"
    assert exe.exists(), "build the CLI with python -m feht_b200.build"
    r = subprocess.run([str(exe), "-i", str(DATA / "test_metadata.txt"), "-d", str(DATA / "test_binary.txt"),
                        "--one", "group A", "--two", "group B", "-f", "0.6", "-c", "none"],
                       capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr
    want = capi.host_run((DATA / "test_metadata.txt").read_bytes(), (DATA / "test_binary.txt").read_bytes(),
                         "group A", "group B", correction="none", ratio_filter=0.6)
    assert r.stdout == want
    assert b"binary4\t0\t5\t3\t1\t4.7619047619047616e-2\t-0.75\n" in r.stdout     # SURVEY.md 8c cross-check
    bad = subprocess.run([str(exe), "-i", "/nonexistent", "-d", str(DATA / "test_binary.txt")], capture_output=True)
    assert bad.returncode != 0


# ------------------------------------------------------------- BASELINE sizes: configs 3, 4, 5
def _check_sampled_tests_against_oracle(ctx, sp_o, snp, comps, vos_list, rows, n_total, correction, rf):
    """Regenerate `rows` (marker-row indices) on the CPU from the counter-based generator and compare the
    GPU's output rows for them with the oracle, comparison by comparison."""
    for ci, (g1, g2) in comps.items():
        got = ctx.fetch(ci)
        pos = {int(r): i for i, r in enumerate(got["row"])}
        for row in rows:
            unit = row // 4 if snp else row
            cells = oracle.synth_fill(sp_o, snp, int(unit), 1)
            if snp:
                cells = snp_expand(cells)[row % 4: row % 4 + 1]
            want = oracle.run_comparison(cells, g1, g2, correction=correction, bonferroni_n=n_total)
            keep = rf <= abs(want["ratio"][0])
            assert (row in pos) == bool(keep), (ci, row, want["ratio"][0])
            if keep:
                i = pos[row]
                assert (got["goa"][i], got["gob"][i], got["gta"][i], got["gtb"][i]) == \
                    (want["goa"][0], want["gob"][0], want["gta"][0], want["gtb"][0]), (ci, row)
                assert got["ratio"][i] == want["ratio"][0]
                l = want["goa"][0] + want["gob"][0] + want["gta"][0] + want["gtb"][0]
                assert p_close(got["p"][i:i + 1], want["p"], np.array([l >= 1000])).all(), (ci, row)


def test_config3_snp_full_size(ctx):
    """Config 3: 500k SNP sites x 10k genomes, one 4-value category -> 10 comparisons x 2M allele rows."""
    S, SITES = 10_000, 500_000
    sp_o = make_spec(oracle.SynthSpec, S, law="uniform_half")
    ctx.set_samples(S)
    ctx.generate_synthetic(copy_spec(sp_o, capi.SynthSpec), True, 0, SITES)
    vos = (np.arange(S) * 4 // S).astype(np.int32)
    ctx.add_category(vos, 4)
    assert ctx.num_rows == 4 * SITES and ctx.num_comparisons == 10
    ctx.run(capi.CORRECTION_BONFERRONI, 0.0)
    assert ctx.result_total() == 10 * 4 * SITES
    groups = [np.flatnonzero(vos == v) for v in range(4)]
    comps = {0: (groups[0], groups[1]), 5: (groups[2], groups[3]),
             8: (groups[2], np.flatnonzero(vos != 2))}
    for ci in comps:
        got = ctx.fetch(ci)
        assert (np.sort(got["row"]) == np.arange(4 * SITES)).all()
        ar = np.abs(got["ratio"])
        assert (ar[:-1] >= ar[1:]).all()
        ties = ar[:-1] == ar[1:]
        assert (got["row"][:-1][ties] < got["row"][1:][ties]).all()
        # the four allele rows of a site partition its valid calls: per group, sum of '+' = N
        o = np.argsort(got["row"])
        n1 = (got["goa"] + got["gob"])[o].reshape(-1, 4)
        assert (n1 == n1[:, :1]).all()
        assert (got["goa"][o].reshape(-1, 4).sum(1) == n1[:, 0]).all()
    rows = np.array([0, 1, 2, 3, 4 * 123456 + 2, 4 * SITES - 1, 777777])
    _check_sampled_tests_against_oracle(ctx, sp_o, True, comps, None, rows, 4 * SITES, 1, 0.0)


def test_config4_category_gemm_full_size(ctx):
    """Config 4: 2M markers x 20k genomes, one 50-value category (1,225 pairs + 50 one-vs-rest), K2 GEMM
    engine, ratioFilter 0.5.  Survivors checked against the oracle; a sample of non-survivors as well."""
    S, R, V = 20_000, 2_000_000, 50
    sp_o = make_spec(oracle.SynthSpec, S)
    ctx.set_samples(S)
    ctx.generate_synthetic(copy_spec(sp_o, capi.SynthSpec), False, 0, R)
    vos = (np.arange(S) // (S // V)).astype(np.int32)
    ctx.add_category(vos, V)
    assert ctx.num_comparisons == V * (V - 1) // 2 + V
    ctx.run(capi.CORRECTION_BONFERRONI, 0.5)
    t = ctx.timings()
    assert t["engine"] == capi.ENGINE_GEMM_FP4              # what AUTO picks for a 50-value category
    groups = [np.flatnonzero(vos == v) for v in range(V)]
    # comparison indices: pair (i,j) -> i*V - i(i+1)/2 + j-i-1 ; one-vs-rest v -> V(V-1)/2 + v
    pidx = lambda i, j: i * V - i * (i + 1) // 2 + (j - i - 1)
    comps = {pidx(0, 25): (groups[0], groups[25]), pidx(24, 25): (groups[24], groups[25]),
             pidx(3, 49): (groups[3], groups[49]), V * (V - 1) // 2 + 7: (groups[7], np.flatnonzero(vos != 7))}
    total = ctx.result_total()
    assert 0 < total < 10_000_000
    for ci in comps:
        got = ctx.fetch(ci)
        ar = np.abs(got["ratio"])
        assert (ar >= 0.5).all() and (ar[:-1] >= ar[1:]).all()
        rows = np.concatenate([got["row"][:6], got["row"][-3:], np.array([0, 5, R - 1])])
        _check_sampled_tests_against_oracle(ctx, sp_o, False, {ci: comps[ci]}, None, np.unique(rows), R, 1, 0.5)
    # both engines give the same survivors (counts are bit-exact either way); K1 on a 200k-row prefix
    ctx.clear_rows()
    ctx.generate_synthetic(copy_spec(sp_o, capi.SynthSpec), False, 0, 200_000)
    res = {}
    for eng in (capi.ENGINE_GEMM, capi.ENGINE_GEMM_FP4, capi.ENGINE_POPC):
        ctx.set_count_engine(eng)
        ctx.run(capi.CORRECTION_BONFERRONI, 0.5, capi.SORT_ABS_RATIO_DESC, R)
        res[eng] = [ctx.fetch(ci) for ci in comps]
    for a, b, c in zip(res[capi.ENGINE_GEMM], res[capi.ENGINE_POPC], res[capi.ENGINE_GEMM_FP4]):
        for k in ("row", "goa", "gob", "gta", "gtb", "ratio", "p"):
            assert (a[k] == b[k]).all() and (a[k] == c[k]).all(), k


def test_config5_shard_four_categories(ctx):
    """Config 5, one GPU's shard of the 8-GPU run: 1.25M of the 10M markers x 50k genomes, four categories
    (2, 5, 10, 20 values -> 281 comparisons), ratioFilter 0.8, Bonferroni n = the global 10M."""
    S, R, TOTAL = 50_000, 1_250_000, 10_000_000
    sp_o = make_spec(oracle.SynthSpec, S)
    ctx.set_samples(S)
    first = 3 * R                                        # the fourth shard of eight
    ctx.set_row_base(first)
    ctx.generate_synthetic(copy_spec(sp_o, capi.SynthSpec), False, first, R)
    rng = np.random.default_rng(5)
    cats = [rng.integers(0, v, S).astype(np.int32) for v in (2, 5, 10, 20)]
    for vos, v in zip(cats, (2, 5, 10, 20)):
        ctx.add_category(vos, v)
    assert ctx.num_comparisons == 1 + 15 + 55 + 210
    ctx.run(capi.CORRECTION_BONFERRONI, 0.8, capi.SORT_ABS_RATIO_DESC, TOTAL)
    assert ctx.timings()["engine"] == capi.ENGINE_GEMM_FP4
    assert ctx.result_total() == 0                       # random categories: no |ratio| >= 0.8
    ctx.run(capi.CORRECTION_BONFERRONI, 0.015, capi.SORT_ABS_RATIO_DESC, TOTAL)
    # comparison 0 = the 2-value category; 1 + 15 + 3 = pair (0,3)... of the 10-value one
    comps = {0: (np.flatnonzero(cats[0] == 0), np.flatnonzero(cats[0] == 1)),
             1 + 4: (np.flatnonzero(cats[1] == 1), np.flatnonzero(cats[1] == 2)),
             1 + 10 + 2: (np.flatnonzero(cats[1] == 2), np.flatnonzero(cats[1] != 2))}
    for ci, (g1, g2) in comps.items():
        got = ctx.fetch(ci)
        assert len(got["row"]) > 0 and got["row"].min() >= first and got["row"].max() < first + R
        for i in list(range(3)) + [len(got["row"]) - 1]:
            row = int(got["row"][i])
            cells = oracle.synth_fill(sp_o, False, row, 1)
            want = oracle.run_comparison(cells, g1, g2, correction=1, bonferroni_n=TOTAL)
            assert (got["goa"][i], got["gob"][i], got["gta"][i], got["gtb"][i]) == \
                (want["goa"][0], want["gob"][0], want["gta"][0], want["gtb"][0]), (ci, row)
            assert got["ratio"][i] == want["ratio"][0] and abs(want["ratio"][0]) >= 0.015
            assert p_close(got["p"][i:i + 1], want["p"], np.array([True])).all()


# ------------------------------------------------------------------------------------ async run / sort engines
def test_run_async_back_to_back_equals_blocking_run(ctx):
    """feht_run_async enqueues without waiting; three runs back to back (the third re-uses the event set of
    the first) must leave the results of a blocking run, and the accumulated stage times must count them."""
    rng = np.random.default_rng(2024)
    cells = _random_chars(rng, 5000, 400)
    g1, g2 = np.arange(0, 150), np.arange(150, 400)
    ctx.set_samples(400)
    ctx.add_rows_chars(cells)
    ctx.add_comparison(g1, g2)
    ctx.run(capi.CORRECTION_BONFERRONI, 0.0)
    want = ctx.fetch(0)
    ctx.timings_accumulated(reset=True)
    for _ in range(3):
        ctx.run_async(capi.CORRECTION_BONFERRONI, 0.0)
    got = ctx.fetch(0)
    for k in want:
        a, b = want[k], got[k]
        assert ((a == b) | ((a != a) & (b != b))).all(), k
    acc = ctx.timings_accumulated()
    assert int(acc["runs"]) == 3 and acc["count_ms"] > 0 and acc["total_ms"] >= acc["count_ms"]
    assert int(acc["launches"]) == 3 * int(ctx.timings()["launches"])
    # a filtered run through the async entry point (it has to wait for the survivor counts itself)
    ctx.run(capi.CORRECTION_NONE, 0.05)
    want = ctx.fetch(0)
    ctx.run_async(capi.CORRECTION_NONE, 0.05)
    got = ctx.fetch(0)
    assert len(want["row"]) > 0 and (want["row"] == got["row"]).all() and (want["p"] == got["p"]).all()


def test_sort_engines_agree(cuda_lib):
    """K4 has two engines: one cooperative launch (one chunk per SM, or several chunks per CTA) and one kernel per pass
    (tile per CTA, look-back).  FEHT_SORT=onesweep forces the second at any size (read once per process,
    hence the subprocesses); both must give the oracle's order on many comparisons, filtered output (name-rank
    passes) and the p-value key."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import oracle, hashlib
        from feht_b200 import capi
        rng = np.random.default_rng(31)
        S, R = 240, 70000        # 21 comparisons x 70k rows = 1.47M dense results: past one chunk per SM
        f = rng.beta(0.4, 0.4, size=(R, 1))
        cells = np.where(rng.random((R, S)) < f, ord('1'), ord('0')).astype(np.uint8)
        cells[rng.random((R, S)) < 0.03] = ord('-')
        vos = rng.integers(0, 6, size=S).astype(np.int32)
        rank = rng.permutation(R).astype(np.int32)
        h = hashlib.sha256()
        with capi.Context(0) as ctx:
            ctx.set_samples(S); ctx.add_rows_chars(cells, name_rank=rank); ctx.add_category(vos, 6)
            for rf, key in ((0.0, capi.SORT_ABS_RATIO_DESC), (0.25, capi.SORT_ABS_RATIO_DESC), (0.0, capi.SORT_PVALUE_ASC)):
                ctx.run(capi.CORRECTION_BONFERRONI, rf, key)
                for ci in range(ctx.num_comparisons):
                    got = ctx.fetch(ci)
                    for k in ("row", "goa", "gob", "gta", "gtb", "p", "ratio"):
                        h.update(np.ascontiguousarray(got[k]).tobytes())
                    r = np.abs(got["ratio"])
                    if key == capi.SORT_ABS_RATIO_DESC:
                        assert (r[:-1] >= r[1:]).all()
                        tie = r[:-1] == r[1:]
                        assert (rank[got["row"][:-1]][tie] < rank[got["row"][1:]][tie]).all()
        print(h.hexdigest())
    """) % (str(ROOT_DIR), str(ROOT_DIR / "tests"))
    digests = {}
    # the cooperative engine's multi-chunk kernel comes in two shapes (two 512-thread CTAs per SM; FEHT_SORT_MULTI=1024:
    # one 1024-thread CTA per SM)
    for name, eng, multi in (("coop", "coop", ""), ("coop1024", "coop", "1024"), ("onesweep", "onesweep", "")):
        env = dict(os.environ, FEHT_SORT=eng, FEHT_SORT_MULTI=multi)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        digests[name] = r.stdout.strip().split()[-1]
    assert digests["coop"] == digests["onesweep"]
    assert digests["coop1024"] == digests["onesweep"]


def test_host_transcoder_planes_equal_gpu_pack(cuda_lib):
    """A chars upload of 64 MB and more is split between the DMA + K0 path and host threads that pack pieces
    with AVX2 (host_pack.cpp).  Both must produce the same bit-planes: ragged rows, rows longer than S, stray
    characters, binary and SNP."""
    rng = np.random.default_rng(99)
    S, R = 2100, 72000
    lens = rng.integers(S - 70, S + 40, size=R)
    lens[:100] = rng.integers(0, 64, size=100)                      # very short rows too
    alphabet = np.frombuffer(b"01-1 0NX", dtype=np.uint8)
    rows = [alphabet[rng.integers(0, len(alphabet), size=int(n))] for n in lens]
    snp_alphabet = np.frombuffer(b"ACGTacgt-N?.x", dtype=np.uint8)
    snp_rows = [snp_alphabet[rng.integers(0, len(snp_alphabet), size=int(n))] for n in lens]
    for snp, data in ((False, rows), (True, snp_rows)):
        planes = {}
        for threads in (0, 5):
            with capi.Context(0) as c:
                c.set_samples(S)
                c.set_host_threads(threads)
                (c.add_snp_chars if snp else c.add_rows_chars)(data)
                assert (c.last_upload_host_rows > 0) == (threads > 0), (snp, threads, c.last_upload_host_rows)
                planes[threads] = [c.read_plane(p, 0, R) for p in range(3 if snp else 2)]
        for a, b in zip(planes[0], planes[5]):
            assert (a == b).all(), snp
        # appended in two calls (the second upload starts at a row offset inside the planes)
        with capi.Context(0) as c:
            c.set_samples(S)
            c.set_host_threads(5)
            add = c.add_snp_chars if snp else c.add_rows_chars
            add(data[: R // 2 + 7])
            add(data[R // 2 + 7:])
            assert c.num_rows == (4 * R if snp else R)
            for p, a in enumerate(planes[0]):
                assert (c.read_plane(p, 0, R) == a).all(), (snp, p)
    # a literal '0' / '1' in SNP rows is rejected by whichever producer meets it: the DMA + K0 path takes the front of
    # the row range, the host threads the back (Table.hs:194-205)
    for where, ch in ((200, b"1"), (R - 3, b"0")):
        bad = list(snp_rows)
        r = bytearray(bytes(bad[where]))
        r[len(r) // 2] = ch[0]
        bad[where] = np.frombuffer(bytes(r), dtype=np.uint8)
        with capi.Context(0) as c:
            c.set_samples(S)
            c.set_host_threads(5)
            with pytest.raises(capi.FehtError):
                c.add_snp_chars(bad)
            assert c.num_rows == 0


def test_filtered_output_single_pass_and_capacity_overflow(ctx):
    """Filtered runs write the survivors in ONE screen pass through a global cursor into arrays sized from a
    guess (1M slots the first time); more survivors than slots makes the library repeat the pass with the exact
    size.  2.1M tests with a ratio filter that nearly all pass (overflow on the first run), then a stricter
    filter in the same context (arrays now larger than needed); both against the oracle."""
    rng = np.random.default_rng(4242)
    S, R = 70, 100000
    cells = _random_chars(rng, R, S, p_missing=0.02)
    vos = rng.integers(0, 6, size=S).astype(np.int32)
    groups = [np.flatnonzero(vos == v) for v in range(6)]
    comps = [(groups[a], groups[b]) for a, b in itertools.combinations(range(6), 2)]
    comps += [(groups[v], np.flatnonzero(vos != v)) for v in range(6)]
    ctx.set_samples(S)
    ctx.add_rows_chars(cells)
    ctx.add_category(vos, 6)
    assert ctx.num_comparisons == len(comps) == 21
    for rf in (1e-9, 0.4):
        ctx.run(capi.CORRECTION_BONFERRONI, rf)
        total = ctx.result_total()
        assert (total > (1 << 20)) == (rf < 0.1), (rf, total)
        for ci in (0, 7, 14, 15, 20):
            g1, g2 = comps[ci]
            want = oracle.run_comparison(cells, g1, g2, correction=1, bonferroni_n=R, n_threads=8)
            order = oracle.filter_and_order(want["ratio"], None, rf)
            got = ctx.fetch(ci)
            assert len(got["row"]) == len(order) == ctx.result_count(ci), (rf, ci)
            assert (got["row"] == order).all(), (rf, ci)
            assert (got["ratio"] == want["ratio"][order]).all()
            for k in ("goa", "gob", "gta", "gtb"):
                assert (got[k] == want[k][order]).all(), (rf, ci, k)
            assert p_close(got["p"], want["p"][order], np.zeros(len(order), bool)).all()


def test_two_devices_in_one_process(cuda_lib):
    """One host process driving a context per GPU (multi-GPU note): each device needs its own
    kernel attributes (dynamic shared memory of the cooperative sort and the GEMM) -- rows split over devices 0 and
    1, host k-way merge, against the oracle.  Skipped on a single-GPU box."""
    import ctypes
    try:
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        pytest.skip("libcudart not loadable by name")
    n = ctypes.c_int(0)
    cudart.cudaGetDeviceCount(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(202)
    S, R = 900, 6000
    cells = _random_chars(rng, R, S)
    vos = rng.integers(0, 12, size=S).astype(np.int32)
    g1, g2 = np.flatnonzero(vos == 3), np.flatnonzero(vos == 7)
    ci = None
    shards = []
    for dev, (lo_, hi_) in enumerate(((0, 2500), (2500, R))):
        with capi.Context(dev) as c:
            c.set_samples(S)
            c.set_row_base(lo_)
            c.add_rows_chars(cells[lo_:hi_])
            c.add_category(vos, 12)
            c.set_count_engine(capi.ENGINE_GEMM)
            c.run(capi.CORRECTION_BONFERRONI, 0.0, capi.SORT_ABS_RATIO_DESC, R)
            assert c.timings()["engine"] == capi.ENGINE_GEMM
            if ci is None:
                ci = next(i for i in range(c.num_comparisons) if c.comparison_info(i)[2:4] == (3, 7))
            shards.append(c.fetch(ci))
    want = oracle.run_comparison(cells, g1, g2, correction=1, bonferroni_n=R)
    order = oracle.filter_and_order(want["ratio"], None, 0.0)
    out_s, out_i = capi.merge_sorted([s["ratio"] for s in shards], [s["name_rank"] for s in shards])
    merged_rows = np.array([shards[s]["row"][i] for s, i in zip(out_s, out_i)])
    assert (merged_rows == order).all()


def test_randomised_shapes_against_oracle(cuda_lib):
    """Forty random jobs -- 1..300 samples, 1..500 rows, binary or SNP, explicit comparisons (disjoint, overlapping,
    empty, single-column groups) or a whole category, either engine, with / without name ranks, both corrections,
    dense and filtered -- each checked row by row against the oracle."""
    import os
    rng = np.random.default_rng(int(os.environ.get("FEHT_FUZZ_SEED", "20261020")))
    for case in range(int(os.environ.get("FEHT_FUZZ_CASES", "40"))):
        S = int(rng.choice([1, 2, 7, 63, 64, 65, 127, 128, 129, int(rng.integers(1, 301))]))
        R = int(rng.choice([1, 2, 31, 32, 33, int(rng.integers(1, 501))]))
        if rng.random() < 0.15:      # table totals past 1000 (chi-square branch), several tiles of every kernel
            S, R = int(rng.integers(1000, 2300)), int(rng.integers(1000, 5000))
        snp = bool(rng.random() < 0.3)
        if snp:
            alphabet = np.frombuffer(b"ACGTacgt-N", dtype=np.uint8)
            sites = rng.choice(alphabet, size=(R, S))
            cells = snp_expand(sites)
        else:
            sites = None
            cells = _random_chars(rng, R, S, p_missing=float(rng.choice([0.0, 0.05, 0.5])))
        n_marker = cells.shape[0]
        use_cat = S >= 4 and rng.random() < 0.4
        comps = []
        if use_cat:
            V = int(rng.integers(2, min(S, 9) + 1))
            vos = rng.integers(-1 if rng.random() < 0.5 else 0, V, size=S).astype(np.int32)
            groups = [np.flatnonzero(vos == v) for v in range(V)]
            comps = [(groups[a], groups[b]) for a, b in itertools.combinations(range(V), 2)]
            if V > 2:
                comps += [(groups[v], np.flatnonzero((vos != v) & (vos >= 0))) for v in range(V)]
        else:
            for _ in range(int(rng.integers(1, 6))):
                kind = rng.integers(0, 4)
                cols = rng.permutation(S)
                if kind == 0:
                    h = max(1, S // 2)
                    comps.append((cols[:h], cols[h:]))                       # second group may be empty (S = 1)
                elif kind == 1:
                    comps.append((cols[: max(1, S // 3)], cols[: max(1, S // 2)]))   # overlapping
                elif kind == 2:
                    comps.append((cols[:1], cols[-1:]))                      # single columns (the same one if S = 1)
                else:
                    comps.append((np.array([], dtype=np.int64), cols))       # empty first group
        rank = rng.permutation(n_marker).astype(np.int32) if rng.random() < 0.5 else None
        correction = int(rng.integers(0, 2))
        rf = float(rng.choice([0.0, 0.0, 0.1, 0.34, 1.0]))
        # one device, or the multi-device handle with two / three shards (on the same GPU when the box has one)
        devices = [0, [0, 0], [0, 0, 0]][int(rng.choice([0, 0, 1, 2]))]
        windowed = len(comps) >= 2 and rng.random() < 0.25
        with capi.Context(devices) as c:
            c.set_samples(S)
            if snp:
                c.add_snp_chars(sites, rank)
            else:
                c.add_rows_chars(cells, rank)
            if use_cat:
                c.add_category(vos, V)
                if not snp and rng.random() < 0.6:
                    c.set_count_engine(int(rng.choice([capi.ENGINE_GEMM, capi.ENGINE_GEMM_FP4])))
            else:
                for g1, g2 in comps:
                    c.add_comparison(g1, g2)
            assert c.num_comparisons == len(comps), case
            what = (f"case {case}: S={S} R={R} snp={snp} cat={use_cat} rf={rf} corr={correction} devices={devices} "
                    f"windowed={windowed} rank={'yes' if rank is not None else 'no'}")
            try:
                if windowed:
                    # the comparisons in two windows (feht_set_comparison_window): each holds its own comparisons only
                    h = int(rng.integers(1, len(comps)))
                    for w0, wn in ((0, h), (h, len(comps) - h)):
                        c.set_comparison_window(w0, wn)
                        c.run(correction, rf)
                        inside = [comps[ci] if w0 <= ci < w0 + wn else None for ci in range(len(comps))]
                        for ci, pair in enumerate(inside):
                            if pair is None:
                                assert len(c.fetch(ci)["row"]) == 0, (ci, "outside the window")
                        _check_against_oracle(c, cells, inside, correction, rf, rank)
                else:
                    c.run(correction, rf)
                    _check_against_oracle(c, cells, comps, correction, rf, rank)
            except AssertionError as e:
                raise AssertionError(f"{what}: {e}") from e
