"""world_size-2 gloo test of the N>1 host path: shard bounds, gather, k-way merge (no GPU).
The per-rank compute is stood in for by the oracle; the merge is the product's feht_merge_sorted."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from feht_b200 import sharding
    rng = np.random.default_rng(42)  # same matrix on every rank
    n_rows, s = 401, 90
    cells = np.where(rng.random((n_rows, s)) < rng.random((n_rows, 1)), ord("1"), ord("0")).astype(np.uint8)
    name_rank = rng.permutation(n_rows).astype(np.int32)
    g1, g2 = np.arange(0, 40), np.arange(40, 90)
    lo, hi = sharding.shard_bounds(n_rows, world)[rank]
    res = oracle.run_comparison(cells[lo:hi], g1, g2, correction=1, bonferroni_n=n_rows)
    order = oracle.filter_and_order(res["ratio"], name_rank[lo:hi], 0.25)
    local = {k: v[order] for k, v in res.items()}
    local["row"] = (order + lo).astype(np.int64)
    local["name_rank"] = name_rank[lo:hi][order]
    merged = sharding.gather_and_merge(local, group=dist.group.WORLD)
    if rank == 0:
        full = oracle.run_comparison(cells, g1, g2, correction=1, bonferroni_n=n_rows)
        want = oracle.filter_and_order(full["ratio"], name_rank, 0.25)
        ok = bool((merged["row"] == want).all() and (merged["p"] == full["p"][want]).all()
                  and (merged["goa"] == full["goa"][want]).all())
        q.put(ok)
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from feht_b200 import sharding
    b = sharding.shard_bounds(10, 4)
    assert b == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert sharding.shard_bounds(3, 8)[-1] == (2, 3)
    assert sum(hi - lo for lo, hi in sharding.shard_bounds(1_000_003, 8)) == 1_000_003


def test_two_rank_gather_and_merge(cuda_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
