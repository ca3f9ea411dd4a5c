#!/usr/bin/env python
"""bench.py -- feht marker-discovery hot path on B200: markers x comparisons tested per second.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  N>1 is launched under torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.

Workload at every N: BASELINE.json configs[1] ("C2"): synthetic pangenome binary matrix,
1,000,000 markers x 5,000 genomes PER GPU (weak scaling: the job has N x 1M marker rows, sharded by
rows, Bonferroni n = the global row count), one --one/--two comparison (genomes 0-2499 vs 2500-4999),
bonferroni, ratioFilter 0 (every test is an output row).  A "step" is one pass of the hot path
(count -> FET/chi-square -> Bonferroni -> ratio filter -> order) over the rank's resident shard.

  value  : whole-job tests/s with the packed matrix resident in HBM (CUDA events, max over ranks)
  e2e    : the same metric through the reference-facing C-ABI calls with HOST buffers: per step
           feht_add_rows_chars (1 byte per call, as the Haskell host holds it; H2D + pack on upload),
           feht_run, feht_result_fetch (D2H of every output row); N>1 adds the gather + k-way merge.
  roofline: the count scan (K1) -- algorithmic bytes / average launch duration vs MEASURED_PEAKS hbm_gbs
  cpu_baseline: the CPU oracle port of the reference algorithm on the box's host cores (bounded sample)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

S = 5000                 # genomes
ROWS_PER_GPU = 1_000_000  # markers per GPU
G1 = np.arange(0, 2500, dtype=np.int32)
G2 = np.arange(2500, 5000, dtype=np.int32)
METRIC = "markers x comparisons tested/sec"
UNIT = "tests/s"


def _spec(cls):
    from helpers import make_spec
    return make_spec(cls, S)


def _k1_traffic():
    """dram bytes read+written by one K1 launch of this workload, from the committed ncu --set full capture."""
    p = ROOT / "profiles" / "k1_traffic.json"
    if p.exists():
        d = json.loads(p.read_text())
        if d.get("rows") == ROWS_PER_GPU and d.get("genomes") == S:
            return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    return None


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md).
    NVML is polled from a thread every ~1 ms (the C-ABI calls release the GIL); a timed region of a
    few milliseconds is far too short for `nvidia-smi -lms`, which is only the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, uuid: str | None = None):
        self.gpu = gpu_index
        self.uuid = uuid
        self.samples = []      # (sm_mhz, reasons bitmask, power_w)
        self.stop_flag = threading.Event()
        self.thread = None
        self.nv = None
        self.handle = None
        self.sm_max = None
        self.proc = None
        self.lines = []

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(self.uuid.encode() if hasattr(self.uuid, "encode") else self.uuid)
                except Exception:
                    h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                idx = self.gpu
                if vis:
                    ent = vis.split(",")[self.gpu].strip()
                    if ent.isdigit():
                        idx = int(ent)
                    else:
                        h = nv.nvmlDeviceGetHandleByUUID(ent.encode())
                if h is None:
                    h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.nv, self.handle = nv, h
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv, self.handle
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((sm, rs, pw))
            except Exception:
                break
            time.sleep(0.001)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.nv is not None:
            self.stop_flag.set()
            self.thread.join(2)
            nv = self.nv
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
                     "hw_power_brake": nv.nvmlClocksEventReasonHwPowerBrakeSlowdown}
            reasons = sorted(n for n, bit in names.items() if any(s[1] & bit for s in self.samples))
            sm = [s[0] for s in self.samples]
            pw = [s[2] for s in self.samples if s[2] is not None]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max,
                    "samples": len(sm), "power_w_max": max(pw) if pw else None, "reasons": reasons,
                    "source": "NVML polled every ~1 ms during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 20"}


def cpu_baseline(sample_rows: int, threads: int, reference_style: bool = True, reps: int = 1):
    """The oracle port of the reference algorithm (per-index char compares, per-term choose) on
    `sample_rows` rows of the same synthetic workload, `reps` times over.  Returns (tests/s, seconds)."""
    import oracle
    sp = _spec(oracle.SynthSpec)
    cells = oracle.synth_fill(sp, False, 0, sample_rows)
    oracle.set_choose_cache(not reference_style)
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.run_comparison(cells, G1, G2, correction=1, bonferroni_n=ROWS_PER_GPU, n_threads=threads)
    dt = time.perf_counter() - t0
    oracle.set_choose_cache(True)
    return sample_rows * reps / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Haskell binary cannot be built
    in this image: no ghc/stack) on all host cores, same workload/metric, each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = 40_000
    import oracle
    sp = _spec(oracle.SynthSpec)
    cells = oracle.synth_fill(sp, False, 0, sample)
    oracle.set_choose_cache(False)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oracle.run_comparison(cells, G1, G2, correction=1, bonferroni_n=ROWS_PER_GPU, n_threads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * args.steps / total
    sample_txt = (f"{sample} of 1,000,000 marker rows x 5000 genomes per step, oracle port of the reference "
                  f"algorithm (4 index walks per test + chi-square/FET), row-sharded over {cores} threads; "
                  f"the real feht binary is single-threaded (no ghc/stack in the image to build it)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 popcount + f64",
        "data": "synthetic",
        "config": {"workload": "C2 sample: synthetic pangenome binary matrix 1M markers x 5k genomes, single "
                               "--one/--two comparison", "sample_rows_per_step": sample, "genomes": S,
                   "comparisons": 1, "correction": "bonferroni", "ratio_filter": 0.0},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def _guard_stdout():
    """NCCL / torchrun chatter goes to fd 1; the driver wants ONE JSON line there.  Point fd 1 at
    stderr for the whole run and return a file object on the real stdout for the final line."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="marker rows per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=60_000)
    args = ap.parse_args()
    if args.impl == "ours" and args.warmup < 3:
        args.warmup = 3          # timing rule: at least three untimed steps (reported as run)
    if args.impl == "reference":
        return run_reference(args)
    real_stdout = _guard_stdout()

    import torch
    import torch.distributed as dist

    from feht_b200 import capi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libfeht_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = None
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    R = args.rows
    total_rows = R * world
    ctx = capi.Context(local_rank)
    # a REAL stream: torch's default stream has handle 0, which feht_set_stream reads as "use the context's own
    # (non-blocking) stream" -- events recorded on the legacy default stream would then not order against the work
    stream = torch.cuda.Stream(device=dev)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    ctx.set_samples(S)
    # host cores that transcode part of a chars upload (e2e only): this rank's share of the box
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_threads = max(1, (os.cpu_count() or 1) // max(1, local_world) - 1)   # one core drives the DMA path
    if os.environ.get("FEHT_BENCH_HOST_THREADS"):
        host_threads = int(os.environ["FEHT_BENCH_HOST_THREADS"])
    ctx.set_host_threads(host_threads)
    ctx.set_row_base(rank * R)
    ctx.reserve_rows(R)
    ctx.generate_synthetic(_spec(capi.SynthSpec), False, rank * R, R)
    ctx.add_comparison(G1, G2)

    # ------------------------------------------------------------------ resident: `value`
    for _ in range(args.warmup):
        ctx.run(capi.CORRECTION_BONFERRONI, 0.0, capi.SORT_ABS_RATIO_DESC, total_rows)
    barrier()
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.timings_accumulated(reset=True)
    barrier()
    ev0.record(stream)
    # steps are enqueued back to back (feht_run_async: nothing in a dense run needs the host); the stage
    # times come from the library's own CUDA events around every launch group, summed over the K steps
    for _ in range(args.steps):
        ctx.run_async(capi.CORRECTION_BONFERRONI, 0.0, capi.SORT_ABS_RATIO_DESC, total_rows)
    ev1.record(stream)
    barrier()
    stage = ctx.timings_accumulated(reset=True)
    assert int(stage["runs"]) == args.steps
    t = ctx.timings()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    # the window [ev0, ev1] sits on the stream the runs were enqueued on, so it must contain every run's own events
    assert float(ms.item()) >= 0.999 * stage["total_ms"], (float(ms.item()), stage)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = total_rows * 1 * args.steps / (ms_total * 1e-3)
    count_bytes = t["count_bytes"]
    count_ms = stage["count_ms"] / max(1.0, stage["count_launches"])
    peak, peak_src = _peaks()
    achieved = count_bytes / (count_ms * 1e-3) / 1e9

    # ------------------------------------------------------------------ end to end: `e2e`
    e2e = None
    if not args.no_e2e:
        # host-side copy of the shard as the reference host holds it: one byte per call
        chars = torch.empty((R, S), dtype=torch.uint8, pin_memory=True)
        offsets = (torch.arange(R + 1, dtype=torch.int64) * S).pin_memory()
        shifts = torch.arange(64, device=dev, dtype=torch.int64)
        chunk = 20_000
        for r0 in range(0, R, chunk):
            n = min(chunk, R - r0)
            v = torch.from_numpy(ctx.read_plane(0, r0, n).view(np.int64)).to(dev)
            d = torch.from_numpy(ctx.read_plane(1, r0, n).view(np.int64)).to(dev)
            vb = ((v.unsqueeze(-1) >> shifts) & 1).reshape(n, -1)[:, :S]
            db = ((d.unsqueeze(-1) >> shifts) & 1).reshape(n, -1)[:, :S]
            c = torch.where(db == 0, torch.full_like(vb, ord("-")), vb + ord("0")).to(torch.uint8)
            chars[r0:r0 + n].copy_(c)
            del v, d, vb, db, c
        torch.cuda.synchronize()
        # result buffers: rank 0 receives the merged rows of the whole job
        RO = total_rows if rank == 0 else 1
        out = {k: torch.empty(RO, dtype=getattr(torch, dt), pin_memory=True)
               for k, dt in sharding.DEV_DTYPES.items()}
        ptrs = {k: v.data_ptr() for k, v in out.items()}
        h2d = R * S + (R + 1) * 8 + (len(G1) + len(G2)) * 4
        d2h = (total_rows if world > 1 else R) * (8 + 4 * 4 + 8 + 8 + 4)   # rank 0 reads the merged rows
        times = []
        parts = {"upload_pack_ms": 0.0, "run_ms": 0.0, "fetch_merge_ms": 0.0}
        for i in range(1 + args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            ctx.clear_rows()
            ctx.clear_comparisons()
            ctx.add_rows_chars_raw(chars.data_ptr(), offsets.data_ptr(), R)
            ctx.add_comparison(G1, G2)
            t1 = time.perf_counter()
            ctx.last_upload_host_rows_chars = ctx.last_upload_host_rows
            ctx.run(capi.CORRECTION_BONFERRONI, 0.0, capi.SORT_ABS_RATIO_DESC, total_rows)
            t2 = time.perf_counter()
            if world == 1:
                ctx.fetch_raw(0, ptrs)
            else:
                merged = sharding.gather_and_merge_device(ctx, 0, capi.SORT_ABS_RATIO_DESC, 0, out)
                assert merged is None or len(merged["row"]) == total_rows
            barrier()
            dt = time.perf_counter() - t0
            if i > 0:
                times.append(dt)
                parts["upload_pack_ms"] += 1e3 * (t1 - t0) / args.e2e_steps
                parts["run_ms"] += 1e3 * (t2 - t1) / args.e2e_steps
                parts["fetch_merge_ms"] += 1e3 * (t0 + dt - t2) / args.e2e_steps
        e2e_s = torch.tensor([sum(times) / len(times)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        host_rows = ctx.last_upload_host_rows_chars
        # bytes that crossed the link: raw chars for the DMA + K0 share, two packed planes for the host share
        h2d = (R - host_rows) * S + host_rows * ctx.row_stride_words * 16 + (R + 1) * 8 + (len(G1) + len(G2)) * 4
        e2e = {"value": total_rows / float(e2e_s.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "host_input_bytes_per_step": R * S,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * float(e2e_s.item()),
               "rank0_breakdown_ms": parts,
               "host_transcoder": {"threads": host_threads, "rows_packed_on_host": ctx.last_upload_host_rows_chars,
                                   "note": "feht_add_rows_chars splits the upload: DMA + GPU pack kernel for one end "
                                           "of the row range, AVX-512 / AVX2 host threads packing to 2 bits/call for the other"},
               "input_gbs_per_gpu": R * S / (parts["upload_pack_ms"] * 1e-3) / 1e9,
               "path": "feht_add_rows_chars(pinned host chars, 1 B/call; DMA+K0 and host transcoder threads) -> feht_run -> "
                       + ("feht_result_fetch" if world == 1 else
                          "feht_result_fetch_device, NCCL gather to rank 0, feht_merge_sorted_device (K5), one D2H"),
               "timer": "host perf_counter between barrier+synchronize (covers H2D, kernels, D2H)"}
        if world == 1:
            # the same calls with the host transcoder switched off: every byte goes over the link raw and is packed
            # by the GPU kernel (K0) -- what the upload costs without any help from the host cores
            ctx.set_host_threads(0)
            gtimes = []
            for i in range(3):
                barrier()
                t0 = time.perf_counter()
                ctx.clear_rows()
                ctx.clear_comparisons()
                ctx.add_rows_chars_raw(chars.data_ptr(), offsets.data_ptr(), R)
                ctx.add_comparison(G1, G2)
                ctx.run(capi.CORRECTION_BONFERRONI, 0.0, capi.SORT_ABS_RATIO_DESC, total_rows)
                ctx.fetch_raw(0, ptrs)
                barrier()
                if i > 0:
                    gtimes.append(time.perf_counter() - t0)
            ctx.set_host_threads(host_threads)
            e2e["gpu_pack_only"] = {"value": total_rows / (sum(gtimes) / len(gtimes)), "unit": UNIT,
                                    "ms_per_step": 1e3 * sum(gtimes) / len(gtimes), "h2d_bytes_per_step": R * S,
                                    "note": "feht_set_host_threads(0): DMA + GPU pack kernel only"}
        # second figure: the host already holds the PACKED planes (what the C++ loader of feht_b200/host
        # produces from the TSV, and the only form configs 2-5 fit in host memory in): feht_add_rows_packed
        if world == 1:
            Wd = ctx.row_stride_words
            pv = torch.empty((R, Wd), dtype=torch.int64, pin_memory=True)
            pd = torch.empty((R, Wd), dtype=torch.int64, pin_memory=True)
            for r0 in range(0, R, 200_000):
                n = min(200_000, R - r0)
                pv[r0:r0 + n].copy_(torch.from_numpy(ctx.read_plane(0, r0, n).view(np.int64)))
                pd[r0:r0 + n].copy_(torch.from_numpy(ctx.read_plane(1, r0, n).view(np.int64)))
            ptimes = []
            for i in range(1 + args.e2e_steps):
                barrier()
                t0 = time.perf_counter()
                ctx.clear_rows()
                ctx.clear_comparisons()
                ctx.add_rows_packed_raw(pv.data_ptr(), pd.data_ptr(), R, Wd)
                ctx.add_comparison(G1, G2)
                ctx.run(capi.CORRECTION_BONFERRONI, 0.0, capi.SORT_ABS_RATIO_DESC, total_rows)
                ctx.fetch_raw(0, ptrs)
                barrier()
                if i > 0:
                    ptimes.append(time.perf_counter() - t0)
            e2e["packed_host_input"] = {
                "value": total_rows / (sum(ptimes) / len(ptimes)), "unit": UNIT,
                "ms_per_step": 1e3 * sum(ptimes) / len(ptimes), "h2d_bytes_per_step": 2 * R * Wd * 8,
                "d2h_bytes_per_step": d2h,
                "path": "feht_add_rows_packed(pinned host bit-planes, 2 bits/call) -> feht_run -> feht_result_fetch"}
            del pv, pd
        if world == 1:
            # third figure: the data file's TEXT (tab separated, 2 bytes per call + an 8-character name per line) parsed
            # and packed on the device (feht_add_rows_text) -- a 200k-line slice of the same matrix
            Rt = min(R, 200_000)
            row_b = 8 + 2 * S + 1
            text = torch.empty((Rt, row_b), dtype=torch.uint8, pin_memory=True)
            tn = text.numpy()
            tn[:, 0] = ord("m")
            tn[:, 1:8] = np.frombuffer("".join(np.char.zfill(np.arange(Rt).astype(str), 7)).encode(), dtype=np.uint8).reshape(Rt, 7)
            tn[:, 8:8 + 2 * S:2] = ord("\t")
            tn[:, 9:9 + 2 * S:2] = chars.numpy()[:Rt]
            tn[:, -1] = ord("\n")
            ttimes = []
            for i in range(3):
                barrier()
                t0 = time.perf_counter()
                ctx.clear_rows()
                ctx.clear_comparisons()
                ctx.add_rows_text((text.data_ptr(), Rt * row_b), "\t", False)
                ctx.add_comparison(G1, G2)
                ctx.run(capi.CORRECTION_BONFERRONI, 0.0, capi.SORT_ABS_RATIO_DESC, Rt)
                ctx.fetch_raw(0, ptrs)
                barrier()
                if i > 0:
                    ttimes.append(time.perf_counter() - t0)
            e2e["text_input"] = {"value": Rt / (sum(ttimes) / len(ttimes)), "unit": UNIT, "rows": Rt,
                                 "ms_per_step": 1e3 * sum(ttimes) / len(ttimes), "h2d_bytes_per_step": Rt * row_b,
                                 "text_gbs": Rt * row_b / (sum(ttimes) / len(ttimes)) / 1e9,
                                 "path": "feht_add_rows_text(pinned host text of the data file body, parsed and packed by "
                                         "text_load.cu) -> feht_run -> feht_result_fetch"}
            del text
        del chars

    # ------------------------------------------------------------------ cpu baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        reps = 8                                     # ~1 s of wall clock = ~15 core-seconds on 16 cores
        v_all, dt_all = cpu_baseline(args.cpu_sample_rows, cores, reps=reps)
        v_one, dt_one = cpu_baseline(max(2000, args.cpu_sample_rows // 20), 1, reps=4)
        cpu = {"value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
               "single_thread_value": v_one,
               "sample": f"{args.cpu_sample_rows} of the 1,000,000 marker rows (same generator, same comparison) x {reps} "
                         f"passes, oracle C port of the reference algorithm row-sharded over {cores} threads "
                         f"({dt_all:.2f} s wall = {dt_all * cores:.0f} core-seconds); single-thread figure on "
                         f"{max(2000, args.cpu_sample_rows // 20)} rows x 4 passes ({dt_one:.2f} s) -- the real feht binary is single-threaded and could not be built "
                         f"(no ghc/stack in the image)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 popcount + f64", "data": "synthetic",
            "config": {"workload": "C2: synthetic pangenome binary matrix 1M markers x 5k genomes per GPU, single "
                                   "--one/--two comparison (2500 vs 2500 genomes)",
                       "markers_per_gpu": R, "genomes": S, "comparisons": 1, "correction": "bonferroni",
                       "ratio_filter": 0.0, "sort_key": "abs(ratio) desc, name rank",
                       "l2": "inputs larger than L2 (1.28 GB of bit-planes per pass vs 126 MB L2)",
                       "parallelism": f"rows sharded over {world} GPU(s), no data-path collective",
                       "stage_ms_per_step": {k: stage[k] / args.steps for k in ("count_ms", "test_ms", "order_ms")}},
            "roofline": {"bound": "hbm", "kernel": "count_popc_kernel (K1)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _k1_traffic() if R == ROWS_PER_GPU else None,
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                           "(profiles/k1_traffic.json, profiles/README.md r1b)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": count_bytes, "avg_launch_ms": count_ms},
            "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(stage["launches"]),
            "clocks": clocks,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
