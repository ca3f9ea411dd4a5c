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


DEV_DTYPES = {"row": "int64", "goa": "int32", "gob": "int32", "gta": "int32", "gtb": "int32", "p": "float64",
              "ratio": "float64", "name_rank": "int32"}


def gather_and_merge_device(ctx, comparison: int, sort_key: int = capi.SORT_ABS_RATIO_DESC, dst: int = 0,
                            host_out: dict | None = None):
    """Multi-process, one GPU per rank (torch.distributed / NCCL is only the transport): every rank
    copies its ordered rows of `comparison` into device tensors (feht_result_fetch_device), the
    shards travel GPU-to-GPU to rank `dst` (NVLink), which merges them ON THE DEVICE with K5
    (feht_merge_sorted_device + feht_scatter_rows_device) and does ONE device-to-host copy of the
    merged rows.  Returns a dict of host tensors on `dst` (into `host_out` if given), None elsewhere."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    n_local = ctx.result_count(comparison)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = n_local
    dist.all_reduce(sizes)
    sizes = [int(x) for x in sizes.tolist()]
    n_max = max(sizes + [1])
    local = {k: torch.empty(n_max, dtype=getattr(torch, dt), device=dev) for k, dt in DEV_DTYPES.items()}
    ctx.fetch_device(comparison, {k: v.data_ptr() for k, v in local.items()})
    gathered = {}
    for k, v in local.items():
        lst = [torch.empty_like(v) for _ in range(world)] if rank == dst else None
        dist.gather(v, lst, dst=dst)
        gathered[k] = lst
    if rank != dst:
        return None
    total = sum(sizes)
    key = "p" if sort_key == capi.SORT_PVALUE_ASC else "ratio"
    dest = [torch.empty(max(n, 1), dtype=torch.int64, device=dev) for n in sizes]
    torch.cuda.current_stream().synchronize()
    ctx.merge_sorted_device(sizes, [t.data_ptr() for t in gathered[key]],
                            [t.data_ptr() for t in gathered["name_rank"]], [t.data_ptr() for t in dest], sort_key)
    out = {}
    for k, dt in DEV_DTYPES.items():
        merged = torch.empty(max(total, 1), dtype=getattr(torch, dt), device=dev)
        for s, n in enumerate(sizes):
            if n:
                ctx.scatter_rows_device(gathered[k][s].data_ptr(), merged.data_ptr(), dest[s].data_ptr(), n,
                                        merged.element_size())
        if host_out is not None and k in host_out:
            host_out[k][:total].copy_(merged[:total], non_blocking=True)
            out[k] = host_out[k][:total]
        else:
            out[k] = merged[:total].cpu()
    torch.cuda.current_stream().synchronize()
    return out


def gather_and_merge_all_device(ctx, sort_key: int = capi.SORT_ABS_RATIO_DESC, dst: int = 0,
                                host_out: dict | None = None, want_device: bool = False, times: dict | None = None):
    """The same exchange for EVERY comparison at once (one process per GPU): each rank hands over the ordered rows of
    all its comparisons (feht_result_fetch_device with comparison = -1: eight device-to-device copies into one packed
    buffer), ONE `dist.gather` moves them over NVLink to `dst`, and ONE call of feht_merge_shards_device (K5b) merges all
    comparisons and scatters all columns.  Returns (columns, seg_off) on `dst` -- host tensors (into `host_out` when
    given; device tensors with want_device) and the merged per-comparison offsets -- and (None, None) elsewhere."""
    import torch
    import torch.distributed as dist

    import time
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    t0 = time.perf_counter()

    def lap(name):
        # host wall clock between device synchronisations: coarse, for the breakdown in bench.py only
        nonlocal t0
        if times is not None:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            times[name] = times.get(name, 0.0) + 1e3 * (t1 - t0)
            t0 = t1
    n_comps = ctx.num_comparisons
    seg = np.zeros(n_comps + 1, dtype=np.int64)
    np.cumsum(ctx.result_counts(), out=seg[1:])
    seg_t = torch.from_numpy(seg).to(dev)
    segs = torch.empty((world, n_comps + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(segs, seg_t)
    segs = segs.cpu().numpy()
    sizes = [int(x) for x in segs[:, -1]]
    # one packed buffer per rank (the eight columns back to back, n_max rows each) and ONE collective: eight padded
    # dist.gather calls -- send / recv pairs to rank 0, one tensor after the other -- took 15 ms for 20M rows on 8 GPUs
    n_max = (max(sizes + [1]) + 15) // 16 * 16
    widths = {k: torch.empty(0, dtype=getattr(torch, dt)).element_size() for k, dt in DEV_DTYPES.items()}
    order = sorted(DEV_DTYPES, key=lambda k: -widths[k])          # 8-byte columns first: every column stays aligned
    offs, o = {}, 0
    for k in order:
        offs[k] = o
        o += widths[k] * n_max
    pack = torch.empty(o, dtype=torch.uint8, device=dev)
    if sizes[rank]:
        ctx.fetch_device(-1, {k: pack.data_ptr() + offs[k] for k in DEV_DTYPES})
    lap("sizes_and_local_copy_ms")
    # ONE gather of the packed buffers to `dst` (grouped send / recv: the root's NVLink ingress is the bound, 7 x 110 MB
    # for 20M rows on 8 GPUs).  An all-gather of the same buffers delivers every shard to every rank and took 2.9 ms
    # there; FEHT_EXCHANGE=allgather keeps it for comparison.
    import os
    if os.environ.get("FEHT_EXCHANGE") == "allgather":
        everything = torch.empty(world * o, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(everything, pack)
    elif rank == dst:
        everything = torch.empty(world * o, dtype=torch.uint8, device=dev)
        dist.gather(pack, [everything[s * o:(s + 1) * o] for s in range(world)], dst=dst)
    else:
        dist.gather(pack, None, dst=dst)
    lap("nccl_gather_ms")
    if rank != dst:
        return None, None
    total = sum(sizes)
    merged = {k: torch.empty(max(total, 1), dtype=getattr(torch, dt), device=dev) for k, dt in DEV_DTYPES.items()}
    out_seg = np.zeros(n_comps + 1, dtype=np.int64)
    torch.cuda.current_stream().synchronize()
    shards = [({k: everything.data_ptr() + s * o + offs[k] for k in DEV_DTYPES}, np.ascontiguousarray(segs[s]))
              for s in range(world)]
    ctx.merge_shards_device(shards, n_comps, {k: v.data_ptr() for k, v in merged.items()}, out_seg, sort_key)
    lap("merge_kernel_ms")
    if want_device:
        return {k: v[:total] for k, v in merged.items()}, out_seg
    out = {}
    for k in DEV_DTYPES:
        if host_out is not None and k in host_out:
            host_out[k][:total].copy_(merged[k][:total], non_blocking=True)
            out[k] = host_out[k][:total]
        else:
            out[k] = merged[k][:total].cpu()
    torch.cuda.current_stream().synchronize()
    return out, out_seg
