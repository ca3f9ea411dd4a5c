"""Strong scaling of the sharded BASELINE configs through ONE multi-device handle (feht_create_multi): the whole job
behind the C ABI -- rows sharded over the listed GPUs, every shard counted / tested / ordered on its own device, the
ordered shards copied to the first GPU (cudaMemcpyPeerAsync) and merged there by K5b -- timed by the library's own
events (slowest shard + merge) and by the host clock around feht_run.

    python tools/bench_multi.py c5 --devices 0,1,2,3 [--rows N] [--steps K]

Prints one JSON line per device count 1, 2, 4, ... up to the listed devices."""
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


def config(cfg):
    if cfg == "c3":
        S, snp, rf, rows = 10_000, True, 0.0, 500_000
        cats = [[(np.arange(S) * 4 // S).astype(np.int32), 4]]
        explicit = None
    elif cfg == "c5":
        S, snp, rf, rows = 50_000, False, 0.8, 10_000_000
        rng = np.random.default_rng(5)
        cats = [[rng.integers(0, v, S).astype(np.int32), v] for v in (2, 5, 10, 20)]
        cats[0][0] = (np.arange(S) >= S // 2).astype(np.int32)     # follows the planted split: planted rows survive 0.8
        explicit = None
    elif cfg == "c4":
        S, snp, rf, rows = 20_000, False, 0.5, 2_000_000
        cats = [[(np.arange(S) // 400).astype(np.int32), 50]]
        explicit = None
    elif cfg == "c2":
        S, snp, rf, rows = 5_000, False, 0.0, 8_000_000            # the 8-GPU weak-scaling job as one strong-scaling job
        cats = []
        explicit = (np.arange(0, 2500), np.arange(2500, 5000))
    else:
        raise SystemExit("unknown config")
    return S, snp, rf, rows, cats, explicit


def run(cfg, devices, rows, steps):
    S, snp, rf, rows0, cats, explicit = config(cfg)
    rows = rows or rows0
    with capi.Context(devices if len(devices) > 1 else devices[0]) as ctx:
        ctx.set_samples(S)
        ctx.reserve_rows(rows)
        ctx.generate_synthetic(make_spec(capi.SynthSpec, S, law="uniform_half" if snp else "beta"), snp, 0, rows)
        if explicit is not None:
            ctx.add_comparison(*explicit)
        for vos, nv in cats:
            ctx.add_category(vos, nv)
        n_comps = ctx.num_comparisons
        wall, acc = [], None
        for i in range(2 + steps):
            t0 = time.perf_counter()
            ctx.run(capi.CORRECTION_BONFERRONI, rf)
            wall.append(time.perf_counter() - t0)
            t = ctx.timings()
            if i >= 2:
                acc = t if acc is None else {k: acc[k] + t[k] for k in t}
        avg = {k: acc[k] / steps for k in acc}
        tests = ctx.num_rows * n_comps
        wall_ms = 1e3 * float(np.mean(wall[2:]))
        out = {"config": cfg, "devices": devices, "n_gpus": len(devices), "marker_rows": ctx.num_rows, "samples": S,
               "comparisons": n_comps, "ratio_filter": rf, "merged_rows": ctx.result_total(),
               "slowest_shard_ms": {k: avg[k] for k in ("count_ms", "test_ms", "order_ms")},
               "merge_ms": avg.get("merge_ms", 0.0), "library_total_ms": avg["total_ms"], "host_wall_ms": wall_ms,
               "tests_per_s_device_time": tests / (avg["total_ms"] * 1e-3), "tests_per_s_wall": tests / (wall_ms * 1e-3),
               "path": "feht_create_multi -> feht_generate_synthetic (sharded) -> feht_run (shards side by side, "
                       "cudaMemcpyPeerAsync to device 0, K5b merge) -> merged results"}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cfg")
    ap.add_argument("--devices", default="0")
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    devs = [int(x) for x in a.devices.split(",")]
    n = 1
    while n <= len(devs):
        run(a.cfg, devs[:n], a.rows, a.steps)
        n *= 2
