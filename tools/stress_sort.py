"""Stress the cooperative sort (single- and multi-chunk): many runs, each checked for permutation + order."""
import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from feht_b200 import capi
from helpers import make_spec, copy_spec
import oracle
S = 5000
sp_g = copy_spec(make_spec(oracle.SynthSpec, S), capi.SynthSpec)
bad = 0
for R, reps in ((1_000_000, 120), (1_300_000, 40), (333_333, 60), (2_500_000, 20)):
    with capi.Context(0) as ctx:
        ctx.set_samples(S)
        ctx.generate_synthetic(sp_g, False, 0, R)
        ctx.add_comparison(np.arange(0, 2500), np.arange(2500, 5000))
        ref = None
        for i in range(reps):
            ctx.run_async(capi.CORRECTION_BONFERRONI, 0.0)
            ctx.run_async(capi.CORRECTION_BONFERRONI, 0.0)
            got = ctx.fetch(0)
            r = np.abs(got["ratio"])
            ok = (r[:-1] >= r[1:]).all() and (np.bincount(got["row"], minlength=R) == 1).all()
            if ref is None:
                ref = got["row"].copy()
            ok = ok and (got["row"] == ref).all()
            bad += (not ok)
    print("rows", R, "runs", reps, "bad so far", bad, flush=True)
print("TOTAL BAD", bad)
