"""CPU-side checks of the product library: it loads, exports the whole C ABI, refuses to run
without a B200, and its pure-host pieces (choose table, k-way merge) agree with the oracle."""
import re
from pathlib import Path

import numpy as np
import pytest

import oracle
from feht_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(cuda_lib):
    header = (ROOT / "include" / "feht_cuda.h").read_text()
    declared = set(re.findall(r"\b(feht_[a-z_0-9]+)\s*\(", header))
    declared -= {"feht_ctx"}
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(cuda_lib, name), name
    assert "sm_100a" in capi.version()


def test_choose_table_matches_oracle_bit_for_bit(cuda_lib):
    tab = capi.choose_table()
    assert tab.size == 250500
    i = 0
    bad = 0
    for n in range(1000):
        for k in range(n // 2 + 1):
            if tab[i] != oracle.choose(n, k):
                bad += 1
            i += 1
    assert bad == 0


def test_no_cpu_fallback(cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.FehtError) as e:
        capi.Context(0)
    assert e.value.code == capi.E_NODEVICE
    assert "no CPU fallback" in e.value.msg


def test_merge_sorted_equals_global_order(cuda_lib):
    rng = np.random.default_rng(3)
    n = 5000
    ratio = rng.choice(np.linspace(-1, 1, 41), size=n)  # many ties
    rank = rng.permutation(n).astype(np.int32)
    shard = rng.integers(0, 4, size=n)
    keys, ranks, idxs = [], [], []
    for s in range(4):
        sel = np.flatnonzero(shard == s)
        o = sel[np.lexsort((rank[sel], -np.abs(ratio[sel])))]
        keys.append(ratio[o]); ranks.append(rank[o]); idxs.append(o)
    out_s, out_i = capi.merge_sorted(keys, ranks)
    merged = np.array([idxs[s][i] for s, i in zip(out_s, out_i)])
    want = np.lexsort((rank, -np.abs(ratio)))
    assert (merged == want).all()
    # same order as the oracle's filter_and_order
    assert (oracle.filter_and_order(ratio, rank, 0.0) == want).all()
    # empty shards and p-value key
    out_s, out_i = capi.merge_sorted([np.array([0.1, 0.5]), np.array([]), np.array([0.2])],
                                     [np.array([0, 1]), np.array([]), np.array([2])], capi.SORT_PVALUE_ASC)
    assert list(zip(out_s, out_i)) == [(0, 0), (2, 0), (0, 1)]
