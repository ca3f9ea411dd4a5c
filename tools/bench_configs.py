"""Stage timings of the hot path on the other BASELINE configs (device-resident, CUDA events inside
the library).  python tools/bench_configs.py [c3|c4|c5|c2b] [--rows N] [--engine popc|gemm|auto]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from feht_b200 import capi  # noqa: E402
from helpers import make_spec  # noqa: E402


def run(cfg, rows, engine, steps, ratio_filter=None, samples=None, missing=None):
    eng = {"popc": capi.ENGINE_POPC, "gemm": capi.ENGINE_GEMM, "fp4": capi.ENGINE_GEMM_FP4, "f8": capi.ENGINE_GEMM_F8, "auto": capi.ENGINE_AUTO}[engine]
    with capi.Context(0) as ctx:
        if cfg == "c3":      # SNP 500k sites x 10k, one 4-value category -> 10 comparisons x 2M allele rows
            S, snp, rf = 10_000, True, 0.0
            cats = [(np.arange(S) * 4 // S).astype(np.int32), 4]
            cats = [cats]
        elif cfg == "c4":    # 2M x 20k, one 50-value category -> 1275 comparisons, ratioFilter 0.5
            S, snp, rf = 20_000, False, 0.5
            cats = [[(np.arange(S) // 400).astype(np.int32), 50]]
        elif cfg == "c5":    # 10M x 50k, categories with 2,5,10,20 values -> 281 comparisons, ratioFilter 0.8
            S, snp, rf = 50_000, False, 0.8
            rng = np.random.default_rng(5)
            cats = [[rng.integers(0, v, S).astype(np.int32), v] for v in (2, 5, 10, 20)]
        elif cfg == "c2":    # 1M x 5k, one comparison 2500 vs 2500 (the bench.py workload)
            S, snp, rf = 5_000, False, 0.0
            cats = []
        elif cfg == "c2b":   # 1M x 5k, groups of 300 / 400 genomes: hypergeometric branch (l ~ 693)
            S, snp, rf = 5_000, False, 0.0
            cats = []
        elif cfg.startswith("c2cat"):   # C2 shape with one V-value category counted by K1: c2cat4 -> 2 pair-slots, c2cat8 -> 4, ...
            S, snp, rf = 5_000, False, 0.0
            nv = int(cfg[5:])
            cats = [[(np.arange(S) * nv // S).astype(np.int32), nv]]
        else:
            raise SystemExit("unknown config")
        if ratio_filter is not None:
            rf = ratio_filter
        if samples is not None and cfg == "c3":     # shape experiments (row alignment)
            S = samples
            cats = [[(np.arange(S) * 4 // S).astype(np.int32), 4]]
        ctx.set_samples(S)
        ctx.reserve_rows(rows)
        sp = make_spec(capi.SynthSpec, S, law="uniform_half" if snp else "beta", **({} if missing is None else {"missing": missing}))
        ctx.generate_synthetic(sp, snp, 0, rows)
        if cfg == "c2b":
            ctx.add_comparison(np.arange(0, 300), np.arange(300, 700))
        if cfg == "c2":
            ctx.add_comparison(np.arange(0, 2500), np.arange(2500, 5000))
        for vos, nv in cats:
            ctx.add_category(vos, nv)
        ctx.set_count_engine(eng)
        ncomp = ctx.num_comparisons
        acc = None
        wall = []
        for i in range(1 + steps):
            t0 = time.perf_counter()
            ctx.run(capi.CORRECTION_BONFERRONI, rf)
            wall.append(time.perf_counter() - t0)
            t = ctx.timings()
            if i > 0:
                acc = t if acc is None else {k: acc[k] + t[k] for k in t}
        avg = {k: acc[k] / steps for k in acc}
        tests = ctx.num_rows * ncomp
        out = {"config": cfg, "engine": engine, "engine_used": int(t["engine"]), "marker_rows": ctx.num_rows,
               "samples": S, "comparisons": ncomp, "ratio_filter": rf, "survivors": ctx.result_total(),
               "count_ms": avg["count_ms"], "test_ms": avg["test_ms"], "order_ms": avg["order_ms"],
               "total_ms": avg["total_ms"], "wall_ms": 1e3 * float(np.mean(wall[1:])),
               "tests_per_s_device": tests / (avg["total_ms"] * 1e-3),
               "count_gbs": avg["count_bytes"] / (avg["count_ms"] * 1e-3) / 1e9,
               "launches": avg["launches"]}
        if t["engine"] in (capi.ENGINE_GEMM, capi.ENGINE_GEMM_FP4, capi.ENGINE_GEMM_F8):
            npad = 0
            for vos, nv in cats:
                npad += 2 * ((nv + 1) // 2) + (2 if nv > 2 else 0)
            passes = -(-npad // 64)
            npad_tot = sum(min(64, npad - 64 * p) for p in range(passes))
            npad_tot = sum(-(-min(64, npad - 64 * p) // 16) * 16 for p in range(passes))
            ops = 2.0 * ctx.num_rows * (-(-S // 128) * 128) * npad_tot * 2
            out["gemm_int8_tops"] = ops / (avg["count_ms"] * 1e-3) / 1e12
            out["gemm_n_padded"] = npad_tot
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cfg")
    ap.add_argument("--rows", type=int, default=200_000)
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--ratio-filter", type=float, default=None)
    ap.add_argument("--samples", type=int, default=None)
    ap.add_argument("--missing", type=float, default=None, help="share of missing calls of the synthetic matrix (default 0.01)")
    a = ap.parse_args()
    run(a.cfg, a.rows, a.engine, a.steps, a.ratio_filter, a.samples, a.missing)
