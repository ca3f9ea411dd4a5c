"""GPU-side debugging aid for K2 (count_gemm.cu): compares the GEMM and popcount counting engines
on impulse matrices (one '1' per row) and random matrices, printing where they differ."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feht_b200 import capi  # noqa: E402


def counts_of(ctx, engine, n_comps):
    ctx.set_count_engine(engine)
    ctx.run(capi.CORRECTION_NONE, 0.0)
    out = []
    for ci in range(n_comps):
        g = ctx.fetch(ci)
        o = np.argsort(g["row"])
        out.append(np.stack([g[k][o] for k in ("goa", "gob", "gta", "gtb")], 1))
    return np.stack(out), ctx.timings()


def main():
    rng = np.random.default_rng(0)
    ok_all = True
    for (R, S, V, kind) in [(256, 128, 4, "impulse"), (300, 300, 5, "impulse"), (700, 1500, 9, "random"),
                            (1000, 5000, 50, "random"), (513, 20000, 50, "random"), (400, 2600, 70, "random")]:
        if kind == "impulse":
            cells = np.full((R, S), ord("0"), np.uint8)
            cells[np.arange(R), np.arange(R) % S] = ord("1")
            cells[np.arange(R), (np.arange(R) * 7 + 3) % S] = ord("-")
        else:
            cells = np.where(rng.random((R, S)) < rng.random((R, 1)), ord("1"), ord("0")).astype(np.uint8)
            cells[rng.random((R, S)) < 0.03] = ord("-")
        vos = (np.arange(S) % (V + 1) - 1).astype(np.int32) if kind == "impulse" else rng.integers(-1, V, S).astype(np.int32)
        with capi.Context(0) as ctx:
            ctx.set_samples(S)
            ctx.add_rows_chars(cells)
            ctx.add_comparison(np.arange(0, S // 2), np.arange(S // 2, S))   # explicit slot first
            ctx.add_category(vos, V)
            n = ctx.num_comparisons
            a, ta = counts_of(ctx, capi.ENGINE_POPC, n)
            b, tb = counts_of(ctx, capi.ENGINE_GEMM, n)
            same = (a == b).all()
            ok_all &= bool(same)
            print(f"R={R} S={S} V={V} {kind}: comparisons={n} equal={same} engine={tb['engine']} "
                  f"count_ms popc={ta['count_ms']:.3f} gemm={tb['count_ms']:.3f}", flush=True)
            if not same:
                bad = np.argwhere(a != b)
                print("  first mismatches (comp,row,field):", bad[:8].tolist())
                for c_, r_, _ in bad[:4]:
                    print("   comp", c_, ctx.comparison_info(int(c_)), "row", r_, "popc", a[c_, r_], "gemm", b[c_, r_])
    print("ALL OK" if ok_all else "MISMATCH")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
