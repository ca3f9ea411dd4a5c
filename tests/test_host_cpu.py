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
    header = (ROOT / "include" / "feht_cuda.h").read_text() + (ROOT / "include" / "feht_host.h").read_text()
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


# ------------------------------------------------------------------ C++ host mirror (no GPU needed)
def _plan_text(meta, data, one, two, delim, mode):
    from oracle import frontend
    pd = frontend.parse_data_file(delim, mode, data)
    table = frontend.get_metadata_from_file(delim, pd, meta)
    cl = frontend.get_comparison_list(table, frontend.generate_categories(one, two))
    n_cols = len(data.split(b"\n")[0].replace(b"\r", b"").split(delim)) - 1
    s = f"columns {n_cols} rows {len(pd.gene_vector_map)}\n".encode()
    for c in cl:
        d = c.details if c.details.endswith(b"\n") else c.details + b"\n"
        s += b"comparison\n" + d
        s += b"g1 " + b"".join(b"%d " % x for x in sorted(c.g1)) + b"\n"
        s += b"g2 " + b"".join(b"%d " % x for x in sorted(c.g2)) + b"\n"
    return s


def test_host_show_double_matches_ghc_show(cuda_lib):
    """GHC `show :: Double` (Comparison.hs:180-181): checked against the oracle front-end's restatement
    and against literal strings feht prints (README.md:116-139)."""
    import math
    import numpy as np
    from feht_b200 import capi
    from oracle import frontend
    lit = {1.0: "1.0", 0.0: "0.0", 0.8: "0.8", -0.8: "-0.8", 0.1: "0.1", 0.05: "5.0e-2", 1e7: "1.0e7",
           9999999.0: "9999999.0", 12345678.9: "1.23456789e7", 0.047619047619047616: "4.7619047619047616e-2",
           -0.6666666666666667: "-0.6666666666666667", 1.1131536985421843e-227: "1.1131536985421843e-227",
           5e-324: "5.0e-324", 1.7976931348623157e308: "1.7976931348623157e308", 100.0: "100.0", 123.456: "123.456"}
    for x, want in lit.items():
        assert capi.host_show_double(x) == want, (x, capi.host_show_double(x))
    assert capi.host_show_double(float("nan")) == "NaN"
    assert capi.host_show_double(float("inf")) == "Infinity" and capi.host_show_double(-float("inf")) == "-Infinity"
    assert capi.host_show_double(-0.0) == "-0.0"
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.random(2000), 10.0 ** rng.uniform(-320, 308, 2000), -rng.random(200) * 1e7,
                         rng.integers(0, 10**9, 500) / 7.0])
    for x in xs:
        assert capi.host_show_double(float(x)) == frontend.show_double(float(x)), x


@pytest.mark.parametrize("data_name,meta_name,delim,mode", [
    ("test_binary.txt", "test_metadata.txt", b"\t", "binary"),
    ("test_snps.txt", "test_metadata.txt", b"\t", "snp"),
    ("test_snps.csv", "test_metadata.csv", b",", "snp"),
])
def test_host_planner_equals_frontend(cuda_lib, fixture_bytes, data_name, meta_name, delim, mode):
    """N2/N3: parseDataFile + getMetadataFromFile + getComparisonList of the C++ host mirror against the
    oracle front-end (descriptions byte for byte, column sets) for the reference fixtures."""
    from feht_b200 import capi
    data, meta = fixture_bytes[data_name], fixture_bytes[meta_name]
    for one, two in (("all all", "all all"), ("group A", "group B"), ("group A B", "position up"),
                     ("position sideways", "all"), ("group C", "group A B")):
        got = capi.host_plan(meta, data, one, two, delim, mode)
        assert got == _plan_text(meta, data, one, two, delim, mode), (one, two)


def test_host_errors_are_the_reference_messages(cuda_lib, fixture_bytes):
    from feht_b200 import capi
    data, meta = fixture_bytes["test_binary.txt"], fixture_bytes["test_metadata.txt"]
    cases = [((b"", data), "File is empty"),
             ((meta.split(b"\n")[0] + b"\n", data), "Metadata file missing header information, data, or both"),
             ((meta + b"nobody\tA\tup\n", data), "name was not found in the data file"),
             ((meta.replace(b"\tup", b"", 1), data), "MetaCategories not equal to MeatValues"),
             ((meta, b""), "No names present in data file")]
    for (m, d), msg in cases:
        with pytest.raises(capi.FehtHostError) as e:
            capi.host_plan(m, d)
        assert msg in str(e.value), (msg, str(e.value))
    with pytest.raises(capi.FehtHostError) as e:
        capi.host_plan(meta, data, one="", two="all all")
    assert "No group one category given" in str(e.value)


def test_host_transcoder_equals_numpy_packing(cuda_lib):
    """feht_pack_chars_host (the AVX2 transcoder of host_pack.cpp, or its scalar twin on CPUs without AVX2) against
    the numpy restatement of the packed layout: ragged rows, rows longer than S, stray characters, binary and SNP,
    one thread and several.  No GPU involved."""
    import numpy as np
    from feht_b200 import capi
    from helpers import pack_rows, pack_snp
    rng = np.random.default_rng(123)
    for S in (1, 63, 64, 65, 200, 1000):
        R = 97
        for snp in (False, True):
            alphabet = np.frombuffer(b"ACGTacgt-N?x" if snp else b"01-1 0NX", dtype=np.uint8)
            lens = rng.integers(max(0, S - 70), S + 40, size=R)
            rows = [alphabet[rng.integers(0, len(alphabet), size=int(n))] for n in lens]
            # what the layout means: characters at index >= len(row) or >= S contribute nothing
            dense = np.full((R, S), ord("-"), dtype=np.uint8)
            for i, r in enumerate(rows):
                m = min(len(r), S)
                dense[i, :m] = r[:m]
            want = pack_snp(dense) if snp else pack_rows(dense)
            for threads in (1, 5):
                got = capi.pack_chars_host([bytes(r) for r in rows], S, snp=snp, n_threads=threads)
                assert len(got) == len(want)
                for g, w in zip(got, want):
                    assert (g == w).all(), (S, snp, threads)
    # SNP rows with a literal '0' / '1' (Table.hs:194-205: counted for all four alleles) cannot be packed: loud error
    for S, pos, ch in ((200, 131, b"1"), (64, 63, b"0"), (65, 64, b"1"), (1000, 7, b"0")):
        row = bytearray(b"ACGT" * (S // 4 + 1))[:S]
        row[pos:pos + 1] = ch
        for threads in (1, 3):
            with pytest.raises(capi.FehtError):
                capi.pack_chars_host([bytes(row), b"ACGT"], S, snp=True, n_threads=threads)
        capi.pack_chars_host([bytes(row)], S, snp=False, n_threads=1)   # binary mode: ordinary calls
