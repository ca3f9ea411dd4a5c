"""Throughput of the device text loader (feht_add_rows_text) on a config-2-shaped data file
body (R lines x 5,000 one-character cells, tab separated, 8-character names: 10,009 bytes per line), from pinned
host memory; next to it the same rows through feht_add_rows_chars (1 byte per call, already parsed) and the C++
host parser of feht_b200/host (parse + plan, one thread) on a sample.
python tools/bench_text.py [--rows 200000]"""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feht_b200 import capi  # noqa: E402


def make_text(R, S, seed=1):
    rng = np.random.default_rng(seed)
    row = 8 + 2 * S + 1
    t = torch.empty(R * row, dtype=torch.uint8).pin_memory()
    a = t.numpy().reshape(R, row)
    names = np.char.zfill(np.arange(R).astype(str), 7)
    a[:, 0] = ord("m")
    a[:, 1:8] = np.frombuffer("".join(names).encode(), dtype=np.uint8).reshape(R, 7)
    a[:, 8:8 + 2 * S:2] = ord("\t")
    cells = a[:, 9:9 + 2 * S:2]
    for r0 in range(0, R, 20000):
        blk = cells[r0:r0 + 20000]
        u = rng.random(blk.shape, dtype=np.float32)
        blk[...] = np.where(u < 0.01, ord("-"), np.where(u < 0.5, ord("1"), ord("0")))
    a[:, -1] = ord("\n")
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200_000)
    ap.add_argument("--samples", type=int, default=5000)
    a = ap.parse_args()
    R, S = a.rows, a.samples
    t = make_text(R, S)
    nbytes = t.numel()
    out = {"rows": R, "samples": S, "text_bytes": nbytes}
    with capi.Context(0) as ctx:
        best = None
        for it in range(4):
            ctx.set_samples(S)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            begin, name_len = ctx.add_rows_text((t.data_ptr(), nbytes), "\t", False)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        assert len(begin) == R and (name_len == 8).all()
        out["text_ms"] = best * 1e3
        out["text_gbs"] = nbytes / best / 1e9
        out["calls_per_s_text"] = R * S / best
        planes_text = [ctx.read_plane(k, 0, 1000) for k in range(2)]
        # the same rows as characters (what a host parser would hand over)
        chars = torch.empty(R * S, dtype=torch.uint8).pin_memory()
        chars.numpy().reshape(R, S)[...] = t.numpy().reshape(R, -1)[:, 9:9 + 2 * S:2]
        off = torch.arange(0, (R + 1) * S, S, dtype=torch.int64).pin_memory()
        best = None
        for it in range(4):
            ctx.set_samples(S)
            t0 = time.perf_counter()
            ctx.add_rows_chars_raw(chars.data_ptr(), off.data_ptr(), R)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        out["chars_ms"] = best * 1e3
        out["calls_per_s_chars"] = R * S / best
        planes_chars = [ctx.read_plane(k, 0, 1000) for k in range(2)]
        assert all((x == y).all() for x, y in zip(planes_text, planes_chars))
    # C++ host parser (single thread): header + a sample of the lines + a two-genome metadata file
    n = min(R, 4000)
    header = b"\t" + b"\t".join(b"g%d" % i for i in range(S)) + b"\n"
    sample = header + bytes(t.numpy()[: n * (8 + 2 * S + 1)])
    meta = b"name\tgrp\n" + b"".join(b"g%d\t%s\n" % (i, b"X" if i < S // 2 else b"Y") for i in range(S))
    t0 = time.perf_counter()
    capi.host_plan(meta, sample, "grp X", "grp Y", b"\t", "binary")
    dt = time.perf_counter() - t0
    out["host_parser_sample_lines"] = n
    out["host_parser_gbs"] = len(sample) / dt / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
