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
