"""Soak test of the stage-parallel MMA issue of K2 (three threads accumulate into the same TMEM accumulators): the mxf4
and i8 engines, run after run on the same matrix, must give the popcount engine's rows every time.
python tools/soak_gemm.py [--runs 30] [--rows 300000]"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from feht_b200 import capi  # noqa: E402
from helpers import make_spec  # noqa: E402


def digest(ctx, n_comps):
    d = ctx.fetch_all()
    h = 0
    for k in ("row", "goa", "gob", "gta", "gtb"):
        h ^= int(np.bitwise_xor.reduce(d[k].astype(np.int64) * np.arange(1, len(d[k]) + 1, dtype=np.int64)))
    return len(d["row"]), h


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=30)
    ap.add_argument("--rows", type=int, default=300_000)
    a = ap.parse_args()
    bad = 0
    for S, V in ((20_000, 50), (5_000, 13), (50_000, 20)):
        with capi.Context(0) as ctx:
            ctx.set_samples(S)
            ctx.generate_synthetic(make_spec(capi.SynthSpec, S), False, 0, a.rows)
            ctx.add_category((np.arange(S) * V // S).astype(np.int32), V)
            n_comps = ctx.num_comparisons
            ctx.set_count_engine(capi.ENGINE_POPC)
            ctx.run(capi.CORRECTION_NONE, 0.3)
            want = digest(ctx, n_comps)
            for eng in (capi.ENGINE_GEMM_FP4, capi.ENGINE_GEMM):
                ctx.set_count_engine(eng)
                for r in range(a.runs):
                    ctx.run(capi.CORRECTION_NONE, 0.3)
                    got = digest(ctx, n_comps)
                    if got != want:
                        bad += 1
                        print("MISMATCH", S, V, eng, r, got, want, flush=True)
            print(f"S={S} V={V}: {want[0]} surviving rows, {2 * a.runs} GEMM runs checked, bad so far {bad}", flush=True)
    print("TOTAL BAD", bad)
    sys.exit(1 if bad else 0)
