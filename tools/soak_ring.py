"""Soak of the TMA-fed ring kernels of K1 (count_popc.cu): the same big matrices counted many times, every result
compared with the register-staged kernels' (run in a child process with FEHT_SNP_RING=0).  Guards the hand-back
logic of the ring (profiles/README.md r1r) against intermittent corruption.
python tools/soak_ring.py [--reps 30]"""
import argparse, hashlib, os, subprocess, sys
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))


def digests(reps):
    from feht_b200 import capi
    from helpers import make_spec
    out = []
    for name, S, R, snp, nv in (("snp4", 10_000, 300_000, True, 4), ("bin8", 5_000, 600_000, False, 8),
                                ("bin4", 5_000, 600_000, False, 4), ("bin14", 3_000, 400_000, False, 14)):
        with capi.Context(0) as ctx:
            ctx.set_samples(S)
            ctx.generate_synthetic(make_spec(capi.SynthSpec, S, law="uniform_half" if snp else "beta"), snp, 0, R)
            ctx.add_category((np.arange(S) * nv // S).astype(np.int32), nv)
            ctx.set_count_engine(capi.ENGINE_POPC)
            seen = set()
            for _ in range(reps):
                ctx.run(capi.CORRECTION_NONE, 0.35)
                h = hashlib.sha256()
                for ci in (0, ctx.num_comparisons // 2, ctx.num_comparisons - 1):
                    got = ctx.fetch(ci)
                    for k in ("row", "goa", "gob", "gta", "gtb"):
                        h.update(np.ascontiguousarray(got[k]).tobytes())
                seen.add(h.hexdigest())
            out.append((name, sorted(seen)))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        for name, ds in digests(1):
            print(name, ds[0])
        sys.exit(0)
    env = dict(os.environ, FEHT_SNP_RING="0")
    ref = dict(l.split() for l in subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True,
                                                 text=True, check=True).stdout.strip().splitlines())
    bad = 0
    for name, ds in digests(a.reps):
        ok = ds == [ref[name]]
        bad += not ok
        print(name, "runs", a.reps, "distinct results", len(ds), "equal to the register-staged kernel:", ok, flush=True)
    print("BAD", bad)
    sys.exit(1 if bad else 0)
