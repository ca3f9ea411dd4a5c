"""FP32 accumulation of the kind::mxf4 engine at the largest sample counts: exact?"""
import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from feht_b200 import capi
rng = np.random.default_rng(5)
for S in (1 << 20, 1 << 22, (1 << 24) - 64, 1 << 24):
    R = 260
    W = (S + 63) // 64
    valid = np.full((R, W), np.uint64(0xFFFFFFFFFFFFFFFF))
    value = np.full((R, W), np.uint64(0xFFFFFFFFFFFFFFFF))
    value[R // 2:] = rng.integers(0, 1 << 63, size=(R - R // 2, W), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(R - R // 2, W), dtype=np.uint64)
    if S % 64:
        tail = np.uint64((1 << (S % 64)) - 1)
        valid[:, -1] &= tail; value[:, -1] &= tail
    vos = np.zeros(S, dtype=np.int32); vos[S // 3:] = 1; vos[2 * (S // 3):] = 2       # 3 values -> total column = S
    res = {}
    with capi.Context(0) as ctx:
        ctx.set_samples(S); ctx.add_rows_packed(value, valid); ctx.add_category(vos, 3)
        for eng in (capi.ENGINE_GEMM, capi.ENGINE_GEMM_FP4):
            ctx.set_count_engine(eng)
            ctx.run(capi.CORRECTION_NONE, 0.0)
            res[eng] = [ctx.fetch(ci) for ci in range(ctx.num_comparisons)]
    bad8 = bad4 = 0
    n = [int((vos == v).sum()) for v in range(3)]
    for ci in range(len(res[capi.ENGINE_GEMM])):
        for k in ("row", "goa", "gob", "gta", "gtb"):
            b, c = (res[e][ci][k] for e in (capi.ENGINE_GEMM, capi.ENGINE_GEMM_FP4))
            bad4 += int((b != c).sum())
    # all-ones rows (the first half): comparison 0 = value 0 vs 1, comparison 3 = value 0 vs rest
    g = res[capi.ENGINE_GEMM]
    ones = g[0]["row"] < R // 2
    bad8 += int((g[0]["goa"][ones] != n[0]).sum() + (g[0]["gta"][ones] != n[1]).sum())
    ones = g[3]["row"] < R // 2
    bad8 += int((g[3]["goa"][ones] != n[0]).sum() + (g[3]["gta"][ones] != n[1] + n[2]).sum())
    print("S", S, "max count", int(max(r["goa"].max() + r["gta"].max() for r in res[capi.ENGINE_GEMM])), "i8 mismatches", bad8, "fp4 mismatches", bad4, flush=True)
