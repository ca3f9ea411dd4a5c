"""Row sharding across GPUs and the final k-way merge.

Marker rows are partitioned contiguously; every rank counts, tests, filters and orders its own shard
(the Bonferroni n of Comparison.hs:251 is the GLOBAL row count, a host scalar, so no collective is
needed on the data path).  The only exchange is the gather of the already-sorted survivors to rank 0,
which merges them with the C ABI's feht_merge_sorted.
"""
from __future__ import annotations

import numpy as np

from . import capi


def shard_bounds(n_rows: int, world: int) -> list[tuple[int, int]]:
    """Row r goes to rank r*world // n_rows: contiguous, sizes differ by at most one."""
    cuts = [(n_rows * r) // world for r in range(world + 1)]
    return list(zip(cuts[:-1], cuts[1:]))


FIELDS = ("row", "goa", "gob", "gta", "gtb", "p", "ratio", "name_rank")


def gather_and_merge(local: dict, group=None, sort_key: int = capi.SORT_ABS_RATIO_DESC, dst: int = 0):
    """local: one comparison's fetched arrays of this rank (already in final order).
    Returns the merged dict on rank `dst`, None elsewhere.  `group` must be a CPU-capable
    torch.distributed group (gloo); None = single process."""
    if group is None:
        shards = [local]
        rank = dst
    else:
        import torch.distributed as dist
        rank = dist.get_rank(group)
        world = dist.get_world_size(group)
        shards = [None] * world if rank == dst else None
        dist.gather_object({k: local[k] for k in FIELDS}, shards, dst=dst, group=group)
    if rank != dst:
        return None
    key = "p" if sort_key == capi.SORT_PVALUE_ASC else "ratio"
    out_s, out_i = capi.merge_sorted([s[key] for s in shards], [s["name_rank"] for s in shards], sort_key)
    merged = {}
    offs = np.concatenate([[0], np.cumsum([len(s["row"]) for s in shards])])[:-1]
    flat = (offs[out_s] + out_i).astype(np.int64)
    for k in FIELDS:
        merged[k] = np.concatenate([s[k] for s in shards])[flat]
    return merged
