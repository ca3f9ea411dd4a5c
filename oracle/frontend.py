"""Oracle front-end: a Python restatement of feht's file parsing, comparison planning and output
formatting, so the fixtures of config 1 can be run end to end without Haskell.

TEST INFRASTRUCTURE ONLY (small inputs; pure-Python loops).  Follows, under /root/reference/src:
  Table.hs:267-279   parseDataFile          Table.hs:160-205  row strings, SNP expansion, replaceChar
  Table.hs:80-115    getMetadataFromFile    Table.hs:229-247  filterTable
  Table.hs:251-261   generateCategories     Comparison.hs:271-364  getComparisonList & friends
  Comparison.hs:129-193  formatComparisonResultsAsTable (filter, order, Haskell `show`)
  app/Main.hs:13-31  main
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from decimal import Decimal

import numpy as np

from . import filter_and_order, run_comparison


class FehtHostError(Exception):
    """The reference calls `error`; the message is kept."""


def _lines(bs: bytes) -> list[bytes]:
    # BS.lines: split at '\n', no trailing empty line; then BS.filter ('\r' /=)
    parts = bs.split(b"\n")
    if parts and parts[-1] == b"":
        parts.pop()
    return [p.replace(b"\r", b"") for p in parts]


def replace_char(match: int, cur: int) -> int:
    n = cur - 32 if 97 <= cur <= 122 else cur  # toUpper (ASCII)
    if n == match:
        return ord("1")
    if n in b"ATCG":
        return ord("0")
    return n


@dataclass
class ParsedDataFile:
    name_column_map: dict  # genome name -> column number (last duplicate wins, M.insert)
    gene_vector_map: dict  # marker name -> bytes (the V.Vector Char), insertion = HashMap semantics


def parse_data_file(delim: bytes, mode: str, bs: bytes) -> ParsedDataFile:
    lines = _lines(bs)
    if not lines:
        raise FehtHostError("No names present in data file")
    names = lines[0].split(delim)[1:]  # first header cell dropped (Table.hs:277)
    ncm = {}
    for i, n in enumerate(names):
        ncm[n] = i
    tuples = []  # built by prepending, then folded into the HashMap: first row in file wins
    for ln in lines[1:]:
        cells = ln.split(delim)
        g, rest = cells[0], b"".join(cells[1:])  # Table.hs:169-170: cells concatenated
        if mode == "snp":
            for suffix, ch in ((b"_a", b"A"), (b"_t", b"T"), (b"_c", b"C"), (b"_g", b"G")):
                tuples.append((g + suffix, bytes(replace_char(ch[0], c) for c in rest)))
        else:
            tuples.append((g, rest))
    gvm = {}
    for name, vec in reversed(tuples):  # foldl' over the reversed list, later insert overwrites
        gvm[name] = vec
    # re-key in file order for deterministic iteration (HashMap order is not reproducible)
    ordered = {}
    for name, _ in tuples:
        if name not in ordered:
            ordered[name] = gvm[name]
    return ParsedDataFile(ncm, ordered)


def get_metadata_from_file(delim: bytes, pd: ParsedDataFile, bs: bytes) -> dict:
    """Table = {(genome name, column): {category: value}}"""
    if len(bs) == 0:
        raise FehtHostError("File is empty")
    lines = _lines(bs)
    if len(lines) < 2:
        raise FehtHostError("Metadata file missing header information, data, or both")
    cats = lines[0].split(delim)[1:]
    table = {}
    for ln in lines[1:]:
        cells = ln.split(delim)
        if len(cells) < 2:
            raise FehtHostError("Metadata line missing name, data, or both")
        name, vals = cells[0], cells[1:]
        if name not in pd.name_column_map:
            raise FehtHostError("name was not found in the data file")
        if len(cats) != len(vals):
            raise FehtHostError("MetaCategories not equal to MeatValues")
        table[(name, pd.name_column_map[name])] = dict(zip(cats, vals))
    return table


def filter_table(table: dict, mc: bytes, all_but: bool, mvs: list[bytes]) -> list[int]:
    """Column numbers of the genomes kept (getListOfColumns . filterTable)."""
    cols = []
    for (name, col), mh in table.items():
        if mc not in mh:
            continue
        v = mh[mc]
        keep = all(v != x for x in mvs) if all_but else any(v == x for x in mvs)
        if keep:
            cols.append(col)
    return cols


def generate_categories(one: str, two: str):
    o, t = one.split(), two.split()
    if not o:
        raise FehtHostError("No group one category given")
    if not t:
        raise FehtHostError("No group two category given")
    return o[0].encode(), [x.encode() for x in o[1:]], t[0].encode(), [x.encode() for x in t[1:]]


def description(xc: bytes, xs: list[bytes], yc: bytes, ys: list[bytes]) -> bytes:
    if not xs:
        raise FehtHostError("No group 1 values")
    g1d = b" Group1: " + b",".join(xs) + b"\n"
    if ys:
        g2d = b" Group2: " + b",".join(ys) + b"\n"
    else:
        g2d = b" Group2: !" + b",".join(xs)  # no newline: Comparison.hs:307
    return b"Group1 category: " + xc + g1d + b"Group2 category: " + yc + g2d


@dataclass
class Comparison:
    g1: list
    g2: list
    details: bytes


def get_comparison_list(table: dict, cats) -> list[Comparison]:
    mc1, mxs1, mc2, mxs2 = cats
    if mc1 == b"all":
        value_sets = {}
        for mh in table.values():
            for k, v in mh.items():
                value_sets.setdefault(k, set()).add(v)
        out = []
        for mc in sorted(value_sets):  # canonical category order (HashMap order in the reference)
            mxs = sorted(value_sets[mc])  # Set.toList: ascending
            for i in range(len(mxs)):
                for j in range(i + 1, len(mxs)):
                    out.append(Comparison(filter_table(table, mc, False, [mxs[i]]),
                                          filter_table(table, mc, False, [mxs[j]]),
                                          description(mc, [mxs[i]], mc, [mxs[j]])))
            if len(mxs) > 2:
                for v in mxs:
                    out.append(Comparison(filter_table(table, mc, False, [v]),
                                          filter_table(table, mc, True, [v]),
                                          description(mc, [v], mc, [])))
        return out
    if mc2 == b"all":
        return [Comparison(filter_table(table, mc1, False, mxs1), filter_table(table, mc1, True, mxs1),
                           description(mc1, mxs1, mc1, []))]
    return [Comparison(filter_table(table, mc1, False, mxs1), filter_table(table, mc2, False, mxs2),
                       description(mc1, mxs1, mc2, mxs2))]


# ---- Haskell `show` -------------------------------------------------------------------------
def show_double(x: float) -> str:
    """GHC's `show :: Double -> String` (Numeric.showFloat / floatToDigits base 10)."""
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "-Infinity" if x < 0 else "Infinity"
    if x < 0 or (x == 0 and math.copysign(1.0, x) < 0):
        return "-" + show_double(-x)
    if x == 0:
        return "0.0"
    # shortest digits that round-trip (what floatToDigits 10 x produces): value = 0.d1d2.. * 10^e
    sign, digs, exp = Decimal(repr(float(x))).as_tuple()
    digits = "".join(map(str, digs)).lstrip("0")
    stripped = digits.rstrip("0")
    exp += len(digits) - len(stripped)
    digits = stripped
    e = len(digits) + exp
    if 0 <= e <= 7:  # FFGeneric: exponent form iff e < 0 || e > 7
        if e == 0:
            return "0." + digits
        ds = digits.ljust(e, "0")
        return ds[:e] + "." + (ds[e:] or "0")
    return f"{digits[0]}.{digits[1:] or '0'}e{e - 1}"


def show_int(n: int) -> str:
    return str(int(n))


COLUMN_HEADER = b"Name\tGroupOne (+)\tGroupOne (-)\tGroupTwo (+)\tGroupTwo (-)\tpValue\tRatio\n"


def run_files(metadata: bytes, data: bytes, one="all all", two="all all", delimiter=b"\t", mode="binary",
              correction="bonferroni", ratio_filter=0.0):
    """Main.hs:13-31 on the oracle.  Returns (blocks, structured) where blocks is the list of output
    ByteStrings (one per comparison that has surviving rows, in canonical comparison order) and
    structured the per-comparison dicts (names in final order + arrays)."""
    cats = generate_categories(one, two)
    pd = parse_data_file(delimiter, mode, data)
    table = get_metadata_from_file(delimiter, pd, metadata)
    cl = get_comparison_list(table, cats)
    names = list(pd.gene_vector_map.keys())
    rows = [pd.gene_vector_map[n] for n in names]
    order_by_name = sorted(range(len(names)), key=lambda i: names[i])
    name_rank = np.zeros(len(names), np.int32)
    for rank, i in enumerate(order_by_name):
        name_rank[i] = rank
    blocks, structured = [], []
    for comp in cl:
        res = run_comparison(rows, comp.g1, comp.g2, correction=1 if correction == "bonferroni" else 0,
                             bonferroni_n=len(rows))
        order = filter_and_order(res["ratio"], name_rank, ratio_filter)
        structured.append({"details": comp.details, "g1": comp.g1, "g2": comp.g2,
                           "names": [names[i] for i in order], "rows": order,
                           **{k: v[order] for k, v in res.items()}})
        if len(order) == 0:
            continue  # Comparison.hs:145-146
        body = []
        for i in order:
            body.append(b"\t".join([names[i], show_int(res["goa"][i]).encode(), show_int(res["gob"][i]).encode(),
                                    show_int(res["gta"][i]).encode(), show_int(res["gtb"][i]).encode(),
                                    show_double(res["p"][i]).encode(), show_double(res["ratio"][i]).encode()])
                        + b"\n")
        blocks.append(b"[#-\n" + comp.details + b"---\n" + COLUMN_HEADER + b"".join(body) + b"-#]\n")
    return blocks, structured
