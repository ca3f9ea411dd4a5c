"""Round-2 additions, through the C ABI on a B200: context re-use, rejected SNP input, the multi-device context
(row sharding + device merge behind feht_create_multi), p-value dedup, blocked dense output."""
import numpy as np
import pytest

import oracle
from feht_b200 import capi

from helpers import p_close, snp_expand

pytestmark = pytest.mark.gpu


def test_set_samples_drops_comparisons_and_categories(ctx):
    """feht_set_samples with a larger S after a category was added: the old per-sample value vector (S entries)
    must not be read with the new S (heap out-of-bounds read in the mask build)."""
    rng = np.random.default_rng(5)
    ctx.set_samples(40)
    ctx.add_rows_chars(rng.choice(np.frombuffer(b"01", np.uint8), size=(10, 40)))
    ctx.add_category(rng.integers(0, 3, 40).astype(np.int32), 3)
    ctx.add_comparison([0, 1], [2, 3])
    assert ctx.num_comparisons == 7
    ctx.run()
    ctx.set_samples(5000)
    assert ctx.num_comparisons == 0
    cells = rng.choice(np.frombuffer(b"01-", np.uint8), size=(64, 5000), p=[.45, .45, .1])
    ctx.add_rows_chars(cells)
    g1, g2 = np.arange(0, 2000), np.arange(2000, 5000)
    ctx.add_comparison(g1, g2)
    ctx.run(capi.CORRECTION_NONE, 0.0)
    want = oracle.run_comparison(cells, g1, g2, correction=0)
    order = oracle.filter_and_order(want["ratio"], None, 0.0)
    got = ctx.fetch(0)
    assert (got["row"] == order).all()
    for k in ("goa", "gob", "gta", "gtb", "ratio"):
        assert (got[k] == want[k][order]).all(), k


def test_snp_chars_with_literal_digits_are_rejected(ctx):
    """Table.hs:194-205: replaceChar leaves a literal '0' / '1' alone, so it counts for all four allele rows; the
    2-bit planes cannot hold that, so feht_add_snp_chars must fail loudly instead of calling it missing."""
    rng = np.random.default_rng(6)
    sites = rng.choice(np.frombuffer(b"ACGT-", np.uint8), size=(50, 333))
    ctx.set_samples(333)
    ctx.add_snp_chars(sites)
    assert ctx.num_rows == 200
    bad = sites.copy()
    bad[17, 200] = ord("1")
    with pytest.raises(capi.FehtError) as ei:
        ctx.add_snp_chars(bad)
    assert ei.value.code == capi.E_INVALID and "all four alleles" in ei.value.msg
    assert ctx.num_rows == 200            # nothing appended
    bad[17, 200] = ord("0")
    with pytest.raises(capi.FehtError):
        ctx.add_snp_chars(bad)
    # the documented way: the expanded rows through the mode-agnostic entry point give the reference's counts
    ctx.set_samples(333)
    expanded = snp_expand(bad)
    ctx.add_rows_chars(expanded)
    g1, g2 = np.arange(0, 150), np.arange(150, 333)
    ctx.add_comparison(g1, g2)
    ctx.run(capi.CORRECTION_NONE, 0.0)
    want = oracle.run_comparison(expanded, g1, g2, correction=0)
    got = ctx.fetch(0)
    back = np.argsort(got["row"])
    for k in ("goa", "gob", "gta", "gtb"):
        assert (got[k][back] == want[k]).all(), k
    # site 17's four allele rows all count the literal '0' at column 200 as absent
    assert all(want["gtb"][4 * 17 + j] >= 1 for j in range(4))


# ------------------------------------------------------------------------- multi-device context
def _device_count():
    import ctypes
    try:
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            cudart = ctypes.CDLL("libcudart.so.12")
        except OSError:
            return 1
    n = ctypes.c_int(0)
    cudart.cudaGetDeviceCount(ctypes.byref(n))
    return max(1, n.value)


def _device_lists():
    """Shards on one GPU always (the driver's test box has one); real device lists when the box has more."""
    lists = [[0, 0], [0, 0, 0]]
    n = _device_count()
    if n >= 2:
        lists += [[0, 1], list(range(min(n, 8)))]
    return lists


def _random_chars(rng, n_rows, n_samples, p_missing=0.05):
    freq = rng.beta(0.3, 0.3, size=(n_rows, 1))
    cells = np.where(rng.random((n_rows, n_samples)) < freq, ord("1"), ord("0")).astype(np.uint8)
    miss = rng.random((n_rows, n_samples)) < p_missing
    cells[miss] = ord("-")
    return cells


def _same_results(a, b, n_comps):
    assert a.result_total() == b.result_total()
    for ci in range(n_comps):
        assert a.result_count(ci) == b.result_count(ci), ci
        x, y = a.fetch(ci), b.fetch(ci)
        for k in x:
            if k == "p":
                assert ((x[k] == y[k]) | (np.isnan(x[k]) & np.isnan(y[k]))).all(), (ci, k)
            else:
                assert (x[k] == y[k]).all(), (ci, k)


@pytest.mark.parametrize("devices", _device_lists())
def test_multi_device_context_equals_single_context_and_oracle(cuda_lib, devices):
    """feht_create_multi: the same calls on one handle that shards the rows over `devices` give the rows of a single
    context in the same order -- explicit comparisons + a category, dense and filtered, both sort keys, ranks given
    and default, rows added in one call and in several (the remap path), and a row base."""
    rng = np.random.default_rng(77)
    S, R = 700, 5000
    cells = _random_chars(rng, R, S)
    cells[:40, :] = ord("0")            # many exact ties: the order is decided by the name rank
    vos = rng.integers(-1, 5, size=S).astype(np.int32)
    g1, g2 = np.arange(0, 300), np.arange(250, 700)        # overlapping groups
    rank = rng.permutation(R).astype(np.int32)
    for given_rank, splits, base in ((True, [R], 0), (False, [R], 1000), (False, [1700, 400, 2900], 0),
                                     (True, [10, R - 10], 0)):
        with capi.Context(0) as one, capi.Context(devices) as many:
            assert many.num_devices == len(devices) and one.num_devices == 1
            for c in (one, many):
                c.set_samples(S)
                c.set_row_base(base)
                r0 = 0
                for n in splits:
                    c.add_rows_chars(cells[r0:r0 + n], rank[r0:r0 + n] if given_rank else None)
                    r0 += n
                c.add_comparison(g1, g2)
                c.add_category(vos, 5)
                assert c.num_rows == R and c.num_comparisons == 1 + 10 + 5
            for corr, rf, key in ((capi.CORRECTION_BONFERRONI, 0.0, capi.SORT_ABS_RATIO_DESC),
                                  (capi.CORRECTION_NONE, 0.25, capi.SORT_ABS_RATIO_DESC),
                                  (capi.CORRECTION_BONFERRONI, 0.1, capi.SORT_PVALUE_ASC)):
                one.run(corr, rf, key)
                many.run(corr, rf, key)
                _same_results(one, many, 16)
            # and against the oracle (the explicit comparison, filtered)
            many.run(capi.CORRECTION_BONFERRONI, 0.2)
            want = oracle.run_comparison(cells, g1, g2, correction=1, bonferroni_n=R)
            order = oracle.filter_and_order(want["ratio"], rank if given_rank else None, 0.2)
            got = many.fetch(0)
            assert (got["row"] == order + base).all()
            for k in ("goa", "gob", "gta", "gtb", "ratio"):
                assert (got[k] == want[k][order]).all(), k
            l = (want["goa"] + want["gob"] + want["gta"] + want["gtb"])[order]
            assert p_close(got["p"], want["p"][order], l >= 1000).all()
            t = many.timings()
            assert t["merge_ms"] > 0 and t["total_ms"] >= t["merge_ms"]
            # planes read back through the handle are the single context's
            for pl in (0, 1):
                assert (many.read_plane(pl, 0, R) == one.read_plane(pl, 0, R)).all()


def test_multi_device_snp_text_packed_and_synthetic(cuda_lib):
    """The other upload paths through a multi-device handle: SNP sites (4 marker rows per unit), packed planes, the
    data file's text (cut at line boundaries, line offsets rebased) and the device generator."""
    from helpers import make_spec, pack_rows
    devices = [0, 0, 0] if _device_count() < 3 else [0, 1, 2]
    rng = np.random.default_rng(78)
    S, n_sites = 333, 901
    sites = rng.choice(np.frombuffer(b"ACGTacgt-N", np.uint8), size=(n_sites, S), p=[.3, .2, .2, .2, .02, .02, .02, .02, .01, .01])
    vos = rng.integers(0, 4, size=S).astype(np.int32)
    with capi.Context(0) as one, capi.Context(devices) as many:
        for c in (one, many):
            c.set_samples(S)
            c.add_snp_chars(sites[:400])
            c.add_snp_chars(sites[400:])
            c.add_category(vos, 4)
            c.run(capi.CORRECTION_BONFERRONI, 0.05)
        assert many.num_rows == 4 * n_sites
        _same_results(one, many, 10)
        # packed binary planes
        cells = _random_chars(rng, 1234, S)
        v, d = pack_rows(cells)
        for c in (one, many):
            c.set_samples(S)
            c.add_rows_packed(v, d)
            c.add_category(vos, 4)
            c.run(capi.CORRECTION_NONE, 0.0)
        _same_results(one, many, 10)
        # text of the data file
        lines = [b"gene%05d\t" % i + b"\t".join(bytes([ch]) for ch in cells[i]) for i in range(len(cells))]
        text = b"\r\n".join(lines) + b"\n"
        for c in (one, many):
            c.set_samples(S)
            begin, name_len = c.add_rows_text(text, "\t", False)
            assert len(begin) == len(cells)
            assert all(text[b:b + n] == b"gene%05d" % i for i, (b, n) in enumerate(zip(begin[:50], name_len[:50])))
            c.set_name_rank(np.arange(len(cells))[::-1].astype(np.int32))
            c.add_category(vos, 4)
            c.run(capi.CORRECTION_NONE, 0.1)
        _same_results(one, many, 10)
        # device generator: global row indices must not depend on the sharding
        sp = make_spec(capi.SynthSpec, S)
        for c in (one, many):
            c.set_samples(S)
            c.generate_synthetic(sp, False, 5000, 3000)
            c.add_comparison(np.arange(0, 150), np.arange(150, 333))
            c.run(capi.CORRECTION_BONFERRONI, 0.0)
        assert (many.read_plane(0, 0, 3000) == one.read_plane(0, 0, 3000)).all()
        _same_results(one, many, 1)
        # a failing shard fails the call loudly and leaves no partial upload behind
        bad = sites.copy()
        bad[800, 5] = ord("1")
        many.set_samples(S)
        with pytest.raises(capi.FehtError):
            many.add_snp_chars(bad)
        assert many.num_rows == 0
        with pytest.raises(capi.FehtError):
            many.set_stream(1234)


def test_merge_shards_device_all_comparisons_in_one_call(cuda_lib):
    """feht_merge_shards_device (K5b) as a one-process-per-GPU host uses it: three single-device contexts hold row
    shards, their ordered columns of ALL comparisons are handed over as device buffers, one call merges them."""
    import torch
    rng = np.random.default_rng(79)
    S, R = 400, 3000
    cells = _random_chars(rng, R, S)
    cells[:30] = ord("1")
    vos = rng.integers(0, 6, size=S).astype(np.int32)
    cuts = [0, 900, 1000, R]
    n_comps = 15 + 6
    dev = torch.device("cuda", 0)
    dt = {"row": torch.int64, "goa": torch.int32, "gob": torch.int32, "gta": torch.int32, "gtb": torch.int32,
          "p": torch.float64, "ratio": torch.float64, "name_rank": torch.int32}
    with capi.Context(0) as whole:
        whole.set_samples(S)
        whole.add_rows_chars(cells)
        whole.add_category(vos, 6)
        whole.run(capi.CORRECTION_BONFERRONI, 0.15)
        shards, keep = [], []
        ctxs = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            c = capi.Context(0)
            ctxs.append(c)
            c.set_samples(S)
            c.set_row_base(a)
            c.add_rows_chars(cells[a:b])
            c.add_category(vos, 6)
            c.run(capi.CORRECTION_BONFERRONI, 0.15, capi.SORT_ABS_RATIO_DESC, R)
            seg = np.zeros(n_comps + 1, np.int64)
            for ci in range(n_comps):
                seg[ci + 1] = seg[ci] + c.result_count(ci)
            cols = {k: torch.empty(max(1, int(seg[-1])), dtype=t, device=dev) for k, t in dt.items()}
            for ci in range(n_comps):
                if seg[ci + 1] > seg[ci]:
                    c.fetch_device(ci, {k: v.data_ptr() + int(seg[ci]) * v.element_size() for k, v in cols.items()})
            keep.append(cols)
            shards.append(({k: v.data_ptr() for k, v in cols.items()}, seg))
        total = sum(int(s[1][-1]) for s in shards)
        assert total == whole.result_total()
        out = {k: torch.empty(max(1, total), dtype=t, device=dev) for k, t in dt.items()}
        out_seg = np.zeros(n_comps + 1, np.int64)
        whole_ptrs = {k: v.data_ptr() for k, v in out.items()}
        ctxs[0].merge_shards_device(shards, n_comps, whole_ptrs, out_seg)
        torch.cuda.synchronize()
        for ci in range(n_comps):
            want = whole.fetch(ci)
            assert out_seg[ci + 1] - out_seg[ci] == len(want["row"])
            for k in dt:
                got = out[k][int(out_seg[ci]):int(out_seg[ci + 1])].cpu().numpy()
                if k == "p":
                    assert ((got == want[k]) | (np.isnan(got) & np.isnan(want[k]))).all(), (ci, k)
                else:
                    assert (got == want[k]).all(), (ci, k)
        for c in ctxs:
            c.close()


# ------------------------------------------------------------------------- N4: one evaluation per distinct table
def test_exact_tests_with_repeated_tables_are_deduplicated_bit_exact(ctx):
    """FET.hs:105-112 is a pure function of the 2x2 table: K3b evaluates every distinct table once (hash dedup) and
    copies p to the other tests.  A matrix built from few distinct row patterns (heavy duplication), plus random
    rows, through both orders and with a filter -- bit-exact against the oracle on every row."""
    rng = np.random.default_rng(80)
    S = 640
    patterns = _random_chars(rng, 37, S, p_missing=0.03)
    cells = np.concatenate([patterns[rng.integers(0, 37, size=20000)], _random_chars(rng, 3000, S, p_missing=0.02)])
    cells = cells[rng.permutation(len(cells))]
    g1, g2 = np.arange(0, 280), np.arange(280, 640)
    vos = (np.arange(S) * 3 // S).astype(np.int32)
    ctx.set_samples(S)
    ctx.add_rows_chars(cells)
    ctx.add_comparison(g1, g2)
    ctx.add_category(vos, 3)
    comps = [(g1, g2)] + [(np.flatnonzero(vos == i), np.flatnonzero(vos == j)) for i in range(3) for j in range(i + 1, 3)] \
        + [(np.flatnonzero(vos == v), np.flatnonzero(vos != v)) for v in range(3)]
    for corr, rf, key in ((1, 0.0, capi.SORT_ABS_RATIO_DESC), (0, 0.05, capi.SORT_ABS_RATIO_DESC),
                          (1, 0.0, capi.SORT_PVALUE_ASC)):
        ctx.run(corr, rf, key)
        for ci, (a, b) in enumerate(comps):
            want = oracle.run_comparison(cells, a, b, correction=corr, bonferroni_n=len(cells), n_threads=8)
            got = ctx.fetch(ci)
            keep = np.flatnonzero(rf <= np.abs(want["ratio"]))
            assert len(got["row"]) == len(keep), (ci, rf)
            back = np.argsort(got["row"], kind="stable")
            assert (got["row"][back] == keep).all()
            for k in ("goa", "gob", "gta", "gtb", "ratio"):
                assert (got[k][back] == want[k][keep]).all(), (ci, k)
            l = (want["goa"] + want["gob"] + want["gta"] + want["gtb"])[keep]
            assert (l < 1000).all()
            assert (got["p"][back] == want["p"][keep]).all(), (ci, "p: the exact branch is bit-exact")
            if key == capi.SORT_PVALUE_ASC:
                assert (got["p"][:-1] <= got["p"][1:]).all()


# ------------------------------------------------------------------------- comparison windows (dense output in blocks)
@pytest.mark.parametrize("devices", [0, [0, 0]])
def test_comparison_window_blocks_equal_one_run(cuda_lib, devices):
    """feht_set_comparison_window: the reference's default ratioFilter 0 keeps every test (Comparison.hs:148), so the
    output of a large category does not fit one run; the host walks the comparisons window by window.  The rows of
    every window must be exactly the rows of the one-shot run, the counts are computed by the first window only, and
    dropping the window restores the ordinary run.  Explicit comparisons + two categories, K1 and the GEMM engines."""
    rng = np.random.default_rng(81)
    S, R = 600, 4000
    cells = _random_chars(rng, R, S)
    vos_a = rng.integers(0, 9, size=S).astype(np.int32)
    vos_b = rng.integers(-1, 3, size=S).astype(np.int32)
    for engine in (capi.ENGINE_POPC, capi.ENGINE_GEMM_FP4, capi.ENGINE_GEMM):
        with capi.Context(devices) as c:
            c.set_samples(S)
            c.add_rows_chars(cells)
            c.add_comparison(np.arange(0, 100), np.arange(100, 350))
            c.add_category(vos_a, 9)
            c.add_comparison(np.arange(50, 300), np.arange(20, 40))
            c.add_category(vos_b, 3)
            c.set_count_engine(engine)
            n_comps = c.num_comparisons
            assert n_comps == 1 + 36 + 9 + 1 + 3 + 3
            for rf in (0.0, 0.12):
                c.set_comparison_window(0, -1)
                c.run(capi.CORRECTION_BONFERRONI, rf)
                whole = [c.fetch(ci) for ci in range(n_comps)]
                first_count_ms = None
                for w0, wn in ((0, 1), (1, 20), (21, 26), (47, 1), (48, 100)):
                    c.set_comparison_window(w0, wn)
                    c.run(capi.CORRECTION_BONFERRONI, rf)
                    t = c.timings()
                    if first_count_ms is None:
                        first_count_ms = t["count_ms"]
                        assert t["count_launches"] >= 1
                    else:
                        assert t["count_launches"] == 0          # the counts of the first window are re-used
                    for ci in range(n_comps):
                        got = c.fetch(ci)
                        if w0 <= ci < w0 + wn:
                            assert len(got["row"]) == len(whole[ci]["row"]), (engine, rf, ci)
                            for k in got:
                                eq = (got[k] == whole[ci][k]) | ((k == "p") & np.isnan(got[k].astype(float)) & np.isnan(whole[ci][k].astype(float)))
                                assert eq.all(), (engine, rf, ci, k)
                        else:
                            assert len(got["row"]) == 0
                    assert c.result_total() == sum(len(whole[ci]["row"]) for ci in range(w0, min(n_comps, w0 + wn)))
                # an ordinary run after the windows counts again
                c.set_comparison_window(0, -1)
                c.run(capi.CORRECTION_BONFERRONI, rf)
                assert c.timings()["count_launches"] >= 1
                assert c.result_total() == sum(len(w["row"]) for w in whole)


# ------------------------------------------------------------------------- BASELINE shapes: large contiguous blocks
def test_config2_block_of_100k_rows_against_oracle(ctx):
    """BASELINE configs[1] (1M x 5k, one comparison): 100,000 contiguous rows (and the last 20,000) through the oracle
    -- every 2x2 table, ratio, p and the rows' relative order; chi-square p-values: >= 95% bit-identical."""
    from helpers import check_block_against_oracle, copy_spec, make_spec
    S, R = 5000, 1_000_000
    sp_o = make_spec(oracle.SynthSpec, S)
    ctx.set_samples(S)
    ctx.generate_synthetic(copy_spec(sp_o, capi.SynthSpec), False, 0, R)
    g1, g2 = np.arange(0, 2500), np.arange(2500, 5000)
    ctx.add_comparison(g1, g2)
    ctx.run(capi.CORRECTION_BONFERRONI, 0.0)
    n, exact, chi_n = check_block_against_oracle(ctx, sp_o, False, {0: (g1, g2)}, 300_000, 100_000, R, 1, 0.0)
    n2, e2, c2 = check_block_against_oracle(ctx, sp_o, False, {0: (g1, g2)}, R - 20_000, 20_000, R, 1, 0.0)
    assert n == 100_000 and n2 == 20_000
    assert chi_n + c2 > 100_000 and (exact + e2) >= 0.95 * (chi_n + c2), (exact + e2, chi_n + c2)
    # a filtered run of the same matrix: planted rows survive, the block's survivors are the oracle's
    ctx.run(capi.CORRECTION_BONFERRONI, 0.3)
    check_block_against_oracle(ctx, sp_o, False, {0: (g1, g2)}, 300_000, 100_000, R, 1, 0.3)


def test_config2b_fet_branch_block_against_oracle(ctx):
    """C2b (groups of 300 / 400 genomes: every test on the hypergeometric branch, with the dedup of K3b): 50,000
    contiguous rows bit-exact against the oracle."""
    from helpers import check_block_against_oracle, copy_spec, make_spec
    S, R = 5000, 1_000_000
    sp_o = make_spec(oracle.SynthSpec, S)
    ctx.set_samples(S)
    ctx.generate_synthetic(copy_spec(sp_o, capi.SynthSpec), False, 0, R)
    g1, g2 = np.arange(0, 300), np.arange(300, 700)
    ctx.add_comparison(g1, g2)
    ctx.run(capi.CORRECTION_BONFERRONI, 0.0)
    n, exact, chi_n = check_block_against_oracle(ctx, sp_o, False, {0: (g1, g2)}, 640_000, 50_000, R, 1, 0.0)
    assert n == 50_000 and chi_n == 0


def test_config3_block_of_20k_allele_rows_all_comparisons(ctx):
    """Config 3 (SNP 500k sites x 10k genomes, one 4-value category): 5,000 contiguous sites = 20,000 allele rows x all
    10 comparisons against the oracle."""
    from helpers import check_block_against_oracle, copy_spec, make_spec
    S, SITES = 10_000, 500_000
    sp_o = make_spec(oracle.SynthSpec, S, law="uniform_half")
    ctx.set_samples(S)
    ctx.generate_synthetic(copy_spec(sp_o, capi.SynthSpec), True, 0, SITES)
    vos = (np.arange(S) * 4 // S).astype(np.int32)
    ctx.add_category(vos, 4)
    ctx.run(capi.CORRECTION_BONFERRONI, 0.0)
    groups = [np.flatnonzero(vos == v) for v in range(4)]
    comps = {}
    ci = 0
    for i in range(4):
        for j in range(i + 1, 4):
            comps[ci] = (groups[i], groups[j])
            ci += 1
    for v in range(4):
        comps[ci] = (groups[v], np.flatnonzero(vos != v))
        ci += 1
    n, exact, chi_n = check_block_against_oracle(ctx, sp_o, True, comps, 123_000, 5_000, 4 * SITES, 1, 0.0)
    assert n == 10 * 20_000 and exact >= 0.95 * chi_n


def test_config4_block_of_5k_rows_24_comparisons(ctx):
    """Config 4 (2M x 20k, one 50-value category, GEMM counts, ratioFilter 0.5 and a low filter that keeps most tests):
    5,000 contiguous rows x 24 comparisons (pairs inside and across the planted split, one-vs-rest) against the oracle."""
    from helpers import check_block_against_oracle, copy_spec, make_spec
    S, R, V = 20_000, 2_000_000, 50
    sp_o = make_spec(oracle.SynthSpec, S)
    ctx.set_samples(S)
    ctx.generate_synthetic(copy_spec(sp_o, capi.SynthSpec), False, 0, R)
    vos = (np.arange(S) // (S // V)).astype(np.int32)
    ctx.add_category(vos, V)
    groups = [np.flatnonzero(vos == v) for v in range(V)]
    pidx = lambda i, j: i * V - i * (i + 1) // 2 + (j - i - 1)
    comps = {}
    for i, j in ((0, 1), (0, 25), (0, 49), (3, 30), (7, 8), (10, 40), (12, 37), (24, 25), (24, 26), (23, 49), (25, 26),
                 (30, 31), (33, 48), (40, 41), (44, 45), (47, 49), (1, 26), (2, 27), (5, 45), (48, 49)):
        comps[pidx(i, j)] = (groups[i], groups[j])
    for v in (0, 7, 24, 49):
        comps[V * (V - 1) // 2 + v] = (groups[v], np.flatnonzero(vos != v))
    assert len(comps) == 24
    ctx.run(capi.CORRECTION_BONFERRONI, 0.5)
    assert ctx.timings()["engine"] == capi.ENGINE_GEMM_FP4
    check_block_against_oracle(ctx, sp_o, False, comps, 1_000_000, 5_000, R, 1, 0.5)
    # a filter most tests pass, on a window of comparisons (the dense output of 1,275 comparisons x 2M rows does not
    # fit one run): window = the first 30 comparisons, 60M result rows
    ctx.set_comparison_window(0, 30)
    ctx.run(capi.CORRECTION_BONFERRONI, 0.02)
    sub = {ci: g for ci, g in comps.items() if ci < 30}
    assert len(sub) >= 2
    n, exact, chi_n = check_block_against_oracle(ctx, sp_o, False, sub, 1_000_000, 5_000, R, 1, 0.02)
    assert n == 5_000 * len(sub)


def test_config5_shard_block_of_5k_rows_24_comparisons(ctx):
    """Config 5, the fourth of eight shards (1.25M of the 10M x 50k rows, four categories, 281 comparisons): 5,000
    contiguous rows x 24 comparisons, with a filter low enough to keep rows, Bonferroni n = the global 10M."""
    from helpers import check_block_against_oracle, copy_spec, make_spec
    S, R, TOTAL = 50_000, 1_250_000, 10_000_000
    sp_o = make_spec(oracle.SynthSpec, S)
    first = 3 * R
    ctx.set_samples(S)
    ctx.set_row_base(first)
    ctx.generate_synthetic(copy_spec(sp_o, capi.SynthSpec), False, first, R)
    rng = np.random.default_rng(5)
    nv = (2, 5, 10, 20)
    cats = [rng.integers(0, v, S).astype(np.int32) for v in nv]
    cats[0] = (np.arange(S) >= S // 2).astype(np.int32)          # follows the planted split: planted rows survive 0.8
    for vos, v in zip(cats, nv):
        ctx.add_category(vos, v)
    assert ctx.num_comparisons == 281
    # comparison indices: category k starts after the comparisons of the categories before it
    start = [0, 1, 1 + 15, 1 + 15 + 55]
    comps = {0: (np.flatnonzero(cats[0] == 0), np.flatnonzero(cats[0] == 1))}
    for k in (1, 2, 3):
        V = nv[k]
        pidx = lambda i, j: start[k] + i * V - i * (i + 1) // 2 + (j - i - 1)
        pairs = [(0, 1), (0, V - 1), (1, 2), (V - 2, V - 1), (1, V - 1)]
        for i, j in pairs:
            comps[pidx(i, j)] = (np.flatnonzero(cats[k] == i), np.flatnonzero(cats[k] == j))
        for v in (0, V - 1, V // 2):
            comps[start[k] + V * (V - 1) // 2 + v] = (np.flatnonzero(cats[k] == v), np.flatnonzero((cats[k] != v)))
    assert len(comps) >= 24
    for rf in (0.8, 0.012):
        ctx.run(capi.CORRECTION_BONFERRONI, rf, capi.SORT_ABS_RATIO_DESC, TOTAL)
        assert ctx.timings()["engine"] == capi.ENGINE_GEMM_FP4
        # (check_block works on global row numbers: the shard's rows are first .. first + R)
        n, exact, chi_n = check_block_against_oracle(ctx, sp_o, False, comps, first + 600_000, 5_000, TOTAL, 1, rf)
        assert n == 5_000 * len(comps)
    assert ctx.result_count(0) > 0


def test_multi_device_edge_cases(cuda_lib):
    """Fewer rows than shards (empty shards), no rows, no comparisons, a window on a multi-device handle, one device
    listed once (a 1-shard "multi" handle), and an invalid device id."""
    rng = np.random.default_rng(83)
    S = 90
    cells = _random_chars(rng, 2, S)
    g1, g2 = np.arange(0, 40), np.arange(40, 90)
    with capi.Context(0) as one, capi.Context([0, 0, 0]) as many, capi.Context([0]) as single:
        for c in (one, many, single):
            c.set_samples(S)
            c.run()                                  # nothing to do
            assert c.result_total() == 0
            c.add_comparison(g1, g2)
            c.run()                                  # no rows
            assert c.result_total() == 0 and c.result_count(0) == 0
            c.add_rows_chars(cells)                  # 2 rows over 3 shards: one shard stays empty
            c.add_category((np.arange(S) % 3).astype(np.int32), 3)
            c.run(capi.CORRECTION_BONFERRONI, 0.0)
        assert many.num_devices == 3 and single.num_devices == 1
        _same_results(one, many, 7)
        _same_results(one, single, 7)
        for c in (one, many):
            c.set_comparison_window(2, 3)
            c.run(capi.CORRECTION_BONFERRONI, 0.0)
            assert c.result_total() == 3 * 2
        _same_results(one, many, 7)
        for c in (one, many):
            c.clear_comparisons()
            c.set_comparison_window(0, -1)
            c.run()
            assert c.result_total() == 0 and c.num_comparisons == 0
    with pytest.raises(capi.FehtError):
        capi.Context([0, 99])
    with pytest.raises(capi.FehtError):
        capi.Context([])
