"""Pins the CPU oracle against every golden value the reference holds for the hot path
(SURVEY.md 8c): FETSpec KATs, ComparisonSpec counting cases, README end-to-end rows, plus the
derived values of an independent restatement."""
import math
from itertools import groupby

import numpy as np
import pytest

import oracle
from oracle import frontend


def test_fetspec_known_answers_bit_exact(golden):
    # /root/reference/test/unit/FETSpec.hs:13-100 -- the reference compares with ~== (about 1e-8);
    # the restatement reproduces the printed doubles bit for bit.
    for goa, gob, gta, gtb, mode, want, _src in golden["fetspec"]:
        got = oracle.fet(goa, gob, gta, gtb, oracle.ONE_TAIL if mode == "one" else oracle.TWO_TAIL)
        assert got == want, (goa, gob, gta, gtb, mode, repr(got), repr(want))


def test_comparisonspec_counting(golden):
    # /root/reference/test/unit/ComparisonSpec.hs:43-51
    for vec, ch, idx, want, _src in golden["countspec"]:
        assert oracle.count_char_in_vector_by_indices(vec.encode(), ch, idx) == want


def test_count_equals_list_count_property():
    # ComparisonSpec.hs:52-57: over all indices the count equals the list count
    rng = np.random.default_rng(5)
    for _ in range(50):
        n = int(rng.integers(1, 60))
        v = bytes(rng.choice(list(b"01-xA"), size=n).tolist())
        for ch in "01-":
            assert oracle.count_char_in_vector_by_indices(v, ch, list(range(n))) == v.count(ch.encode())


def test_count_out_of_bounds_is_an_error():
    assert oracle.count_char_in_vector_by_indices(b"0101", "1", [0, 4]) == -1


def test_derived_choose_and_two_tail(golden):
    for n, k, want in golden["derived_choose"]:
        assert oracle.choose(n, k) == want
    for a, b, c, d, want in golden["derived_two_tail"]:
        assert oracle.fet(a, b, c, d) == want
    for x, want in golden["derived_chi2_p"]:
        assert oracle.chi_squared1_compl_cumulative(x) == want
    L = oracle.lib()
    assert L.fo_log_gamma(0.5) == golden["derived_misc"]["logGamma_0.5"]
    assert L.fo_log_gamma(1.5) == golden["derived_misc"]["logGamma_1.5"]
    assert L.fo_log_beta(291.0, 303.0) == golden["derived_misc"]["logBeta_291_303"]


def test_choose_cache_does_not_change_values():
    oracle.set_choose_cache(False)
    a = [oracle.choose(n, k) for n, k in [(10, 3), (120, 50), (999, 499), (500, 77)]]
    oracle.set_choose_cache(True)
    b = [oracle.choose(n, k) for n, k in [(10, 3), (120, 50), (999, 499), (500, 77)]]
    assert a == b


def test_fet_range_properties():
    # FETSpec.hs:102-125: 0 <= p <= 1 (approximately) for arbitrary Ints, including negatives
    rng = np.random.default_rng(11)
    for _ in range(300):
        w, x, y, z = (int(v) for v in rng.integers(-20, 400, size=4))
        p = oracle.fet(w, x, y, z)
        assert p >= 0 or math.isclose(p, 0, abs_tol=1e-9)
        assert p <= 1 or math.isclose(p, 1, rel_tol=1e-9)


def test_ratio_range_and_bonferroni():
    # ComparisonSpec.hs:15-41,60-65
    rng = np.random.default_rng(12)
    for _ in range(300):
        w, x, y, z = (int(v) for v in rng.integers(-5, 100, size=4))
        r = oracle.calculate_ratio(w, x, y, z)
        assert -1 <= r <= 1
    assert oracle.bonferroni(0.5, 54) == 1.0
    assert oracle.bonferroni(0.001, 54) == 0.001 * 54
    assert math.isnan(oracle.bonferroni(float("nan"), 54))


def test_edge_cases_appendix_a4():
    assert oracle.fet(0, 0, 3, 4) == 1.0 and oracle.calculate_ratio(0, 0, 3, 4) == 0.0  # k = 0
    assert oracle.fet(3, 4, 0, 0) == 1.0                                                # gt = 0
    assert oracle.fet(0, 5, 0, 7) == 1.0                                                # m = 0
    assert math.isnan(oracle.fet(0, 600, 0, 700))     # l >= 1000, zero column margin -> chi2 = 0/0
    assert math.isnan(oracle.fet(600, 0, 700, 0))
    # l = 999 vs 1000 switch (FET.hs:105,109)
    assert oracle.fet(250, 249, 260, 240) == 0.5691584133790883
    p1000 = oracle.fet(250, 250, 260, 240)
    assert p1000 == oracle.chi_squared1_compl_cumulative(oracle.chi_square(250, 250, 260, 240))


def _groups(table, cat, val):
    return sorted(col for (_, col), mh in table.items() if mh.get(cat) == val)


def _find(structured, g1desc, g2desc):
    c1, v1 = g1desc.split()
    c2, v2 = g2desc.split()
    want = f"Group1 category: {c1} Group1: {v1}\nGroup2 category: {c2} Group2: {v2}\n".encode()
    hits = [s for s in structured if s["details"] == want]
    assert len(hits) == 1, want
    return hits[0]


def test_readme_rows_config1(golden, fixture_bytes):
    """README.md:116-139,170-180,214-222: published output rows.  Within a tie of |ratio| the
    reference prints in HashMap order, so ties are compared as sets."""
    meta = fixture_bytes["test_metadata.txt"]
    for fname, mode, g1, g2, rf, rows, _src in golden["readme"]:
        blocks, structured = frontend.run_files(meta, fixture_bytes[fname], mode=mode, ratio_filter=rf)
        s = _find(structured, g1, g2)
        got = {n.decode(): (int(a), int(b), int(c), int(d), frontend.show_double(p), frontend.show_double(r))
               for n, a, b, c, d, p, r in zip(s["names"], s["goa"], s["gob"], s["gta"], s["gtb"], s["p"], s["ratio"])}
        for name, a, b, c, d, p, r in rows:
            assert got[name] == (a, b, c, d, p, r), (fname, name, got[name])
        # the published rows are the head of the block: every unlisted row has |ratio| <= the last listed
        last = abs(float(rows[-1][6]))
        order = [n.decode() for n in s["names"]]
        head = set(order[: len(rows)])
        listed = {r[0] for r in rows}
        tail_ok = all(abs(float(got[n][5])) >= last for n in head)
        assert tail_ok
        # rows strictly above the last listed |ratio| must all be among the listed ones
        assert {n for n in order if abs(float(got[n][5])) > last} <= listed
        if rf == 1.0:
            assert order == ["binary21"]


def test_config1_shape_and_raw_rows(golden, fixture_bytes):
    meta = fixture_bytes["test_metadata.txt"]
    blocks, structured = frontend.run_files(meta, fixture_bytes["test_binary.txt"], correction="none")
    assert len(structured) == 16  # group: 3 pairs + 3 OVA; position: 6 pairs + 4 OVA
    assert all(len(s["names"]) == 54 for s in structured)
    for g1, g2, name, a, b, c, d, p, r in golden["derived_c1_raw"]:
        s = _find(structured, g1, g2)
        i = [n.decode() for n in s["names"]].index(name)
        assert (s["goa"][i], s["gob"][i], s["gta"][i], s["gtb"][i]) == (a, b, c, d)
        assert s["p"][i] == p and s["ratio"][i] == r
    # ordering: |ratio| non-increasing, ties by name ascending
    for s in structured:
        keys = [(-abs(r), n) for r, n in zip(s["ratio"], s["names"])]
        assert keys == sorted(keys)
    # the trailing-space cells ("1 ") of binary9/binary27 shift later columns (SURVEY finding 5)
    pd = frontend.parse_data_file(b"\t", "binary", fixture_bytes["test_binary.txt"])
    assert len(pd.gene_vector_map[b"binary9"]) == 13
    assert len(pd.gene_vector_map[b"binary1"]) == 12


def test_snp_expansion_and_block_format(fixture_bytes):
    meta = fixture_bytes["test_metadata.txt"]
    blocks, structured = frontend.run_files(meta, fixture_bytes["test_snps.txt"], mode="snp")
    assert all(len(s["names"]) == 54 * 4 for s in structured)
    assert blocks[0].startswith(b"[#-\nGroup1 category: ") and blocks[0].endswith(b"-#]\n")
    assert frontend.COLUMN_HEADER in blocks[0]
    # one-vs-rest description has no newline before '---' (Comparison.hs:307)
    ova = [b for b in blocks if b" Group2: !" in b]
    assert ova and all(b"Group2: !A---\n" in b or b"---\n" in b for b in ova)
    # csv variants parse with ',' as the delimiter
    blocks_csv, structured_csv = frontend.run_files(fixture_bytes["test_metadata.csv"],
                                                    fixture_bytes["test_snps.csv"], mode="snp", delimiter=b",")
    assert len(structured_csv) == 16


def test_show_double_matches_ghc_examples():
    sd = frontend.show_double
    assert [sd(x) for x in (1.0, 0.75, -0.8, 0.1, 0.05, 1e7, 9999999.0, 0.0, 4.7619047619047616e-2)] == \
        ["1.0", "0.75", "-0.8", "0.1", "5.0e-2", "1.0e7", "9999999.0", "0.0", "4.7619047619047616e-2"]
    assert sd(float("nan")) == "NaN"


def test_synthetic_generator_is_deterministic_and_plausible():
    sp = oracle.SynthSpec()
    sp.seed = 0xFE47
    sp.n_samples = 300
    sp.missing_thr = int(0.01 * 2**32)
    sp.planted_per_million = 0
    sp.split_col = 150
    sp.thr_hi = int(0.9 * 2**32)
    sp.thr_lo = int(0.1 * 2**32)
    for i in range(256):
        sp.freq_thr[i] = int((i + 0.5) / 256 * 2**32)
    a = oracle.synth_fill(sp, False, 1000, 64)
    b = oracle.synth_fill(sp, False, 1032, 32)
    assert (a[32:] == b).all()                       # counter-based: rows are position independent
    assert set(np.unique(a)) <= {ord("0"), ord("1"), ord("-")}
    frac_missing = (a == ord("-")).mean()
    assert 0.002 < frac_missing < 0.03
    s = oracle.synth_fill(sp, True, 0, 32)
    assert set(np.unique(s)) <= set(b"ACGT-")
