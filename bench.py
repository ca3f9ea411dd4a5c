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


def _synth_rows(spec, snp, row0, n_rows, threads):
    """oracle.synth_fill on `threads` Python threads: rows [row0, row0 + n_rows)."""
    from helpers import synth_rows
    return synth_rows(spec, snp, row0, n_rows, threads)


def cpu_baseline(sample_rows: int, threads: int, reference_style: bool = True, reps: int = 1):
    """The oracle port of the reference algorithm (per-index char compares, per-term choose) on
    `sample_rows` rows of the same synthetic workload, `reps` times over.  Returns (tests/s, seconds)."""
    import oracle
    sp = _spec(oracle.SynthSpec)
    cells = _synth_rows(sp, False, 0, sample_rows, os.cpu_count() or 1)
    oracle.set_choose_cache(not reference_style)
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.run_comparison(cells, G1, G2, correction=1, bonferroni_n=ROWS_PER_GPU, n_threads=threads)
    dt = time.perf_counter() - t0
    oracle.set_choose_cache(True)
    return sample_rows * reps / dt, dt


def _config(world: int, rows_per_gpu: int) -> dict:
    """The `config` object of the JSON line: the same in both arms (ours and --impl reference)."""
    return {"workload": "C2: synthetic pangenome binary matrix 1M markers x 5k genomes per GPU, single "
                        "--one/--two comparison (2500 vs 2500 genomes)",
            "markers_per_gpu": rows_per_gpu, "genomes": S, "comparisons": 1, "correction": "bonferroni",
            "ratio_filter": 0.0, "sort_key": "abs(ratio) desc, name rank",
            "l2": "inputs larger than L2 (1.28 GB of bit-planes per pass vs 126 MB L2)",
            "parallelism": f"rows sharded over {world} GPU(s), no data-path collective"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Haskell binary cannot be built in this
    image: no ghc/stack) on all host cores, same workload/metric.  Each step is one pass over ALL 1,000,000 marker
    rows of configs[1] (5 GB of 1-byte calls, the reference's own layout narrowed from 4 B/call)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = args.ref_rows
    import oracle
    sp = _spec(oracle.SynthSpec)
    cells = _synth_rows(sp, False, 0, sample, cores)
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
    whole = "all" if sample == ROWS_PER_GPU else f"{sample} of the"
    sample_txt = (f"{whole} 1,000,000 marker rows x 5000 genomes per step (one GPU's share of the job), oracle port of the "
                  f"reference algorithm (4 index walks per test + chi-square/FET), row-sharded over {cores} threads; "
                  f"the real feht binary is single-threaded (no ghc/stack in the image to build it)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 popcount + f64",
        "data": "synthetic",
        "config": _config(args.gpus, ROWS_PER_GPU),
        "reference_run": {"markers_per_step": sample, "host_threads": cores, "sampled": sample != ROWS_PER_GPU,
                          "note": "the CPU arm processes one GPU's share of the job (1M marker rows) per step"},
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


def _ncu_facts():
    """ncu-derived per-kernel facts of this round's captures (profiles/roofline_ncu.json): DRAM traffic per launch,
    tensor-pipe and FP64-pipe utilisation.  Constants measured under the profiler, quoted beside the live timings."""
    p = ROOT / "profiles" / "roofline_ncu.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return {}
    return {}


def parity_check_c2(cols: dict, total_rows: int, n_blocks: int = 48, block: int = 250) -> dict:
    """The merged output of the whole job (all N GPUs) against the CPU oracle: `n_blocks` blocks of `block` contiguous
    global rows (first and last rows, every shard boundary, seeded random positions) are regenerated with the
    counter-based generator and run through the oracle with the GLOBAL Bonferroni n; counts and ratio must be
    bit-identical, p within the chi-square tolerance, and the rows' relative order in the merged output the oracle's.
    Over ALL rows: the merged rows are a permutation of [0, total_rows) and the keys (|ratio| desc, rank asc) are
    non-increasing."""
    import oracle
    from helpers import p_close
    row = np.asarray(cols["row"])
    ratio = np.asarray(cols["ratio"])
    rank = np.asarray(cols["name_rank"])
    bad = {"not_a_permutation": 0, "order": 0, "counts": 0, "ratio": 0, "p": 0, "position": 0}
    if len(row) != total_rows or (np.bincount(row, minlength=total_rows) != 1).any():
        bad["not_a_permutation"] = 1
        return {"rows": 0, "mismatches": 1, "detail": bad}
    ar = np.abs(ratio)
    bad["order"] = int((ar[:-1] < ar[1:]).sum() + ((ar[:-1] == ar[1:]) & (rank[:-1] >= rank[1:])).sum())
    pos = np.empty(total_rows, dtype=np.int64)
    pos[row] = np.arange(total_rows)
    rng = np.random.default_rng(0xFE47)
    starts = {0, total_rows - block}
    for k in range(1, total_rows // ROWS_PER_GPU):
        starts.add(k * ROWS_PER_GPU - block // 2)          # straddles the boundary between two GPUs' shards
    while len(starts) < n_blocks:
        starts.add(int(rng.integers(0, total_rows - block)))
    sp = _spec(oracle.SynthSpec)
    checked = 0
    s_rows, s_ratio = [], []
    for r0 in sorted(starts):
        cells = oracle.synth_fill(sp, False, r0, block)
        want = oracle.run_comparison(cells, G1, G2, correction=1, bonferroni_n=total_rows, n_threads=4)
        at = pos[r0:r0 + block]
        for k in ("goa", "gob", "gta", "gtb"):
            bad["counts"] += int((np.asarray(cols[k])[at] != want[k]).sum())
        bad["ratio"] += int((ratio[at] != want["ratio"]).sum())
        l = want["goa"] + want["gob"] + want["gta"] + want["gtb"]
        bad["p"] += int((~p_close(np.asarray(cols["p"])[at], want["p"], l >= 1000)).sum())
        s_rows.append(np.arange(r0, r0 + block))
        s_ratio.append(want["ratio"])
        checked += block
    # the sampled rows among themselves: oracle order (filter 0, ties by name rank = global row) == merged order
    s_rows = np.concatenate(s_rows)
    s_ratio = np.concatenate(s_ratio)
    order = oracle.filter_and_order(s_ratio, s_rows.astype(np.int32), 0.0)
    bad["position"] = int((np.argsort(pos[s_rows], kind="stable") != order).sum())
    return {"rows": checked, "blocks": len(starts), "mismatches": int(sum(bad.values())), "detail": bad,
            "checked": "counts/ratio bit-exact, p (chi-square tolerance), relative order of the sampled rows vs the "
                       "oracle with the global Bonferroni n; all rows: permutation + keys non-increasing"}


# ------------------------------------------------------------------ the other BASELINE configs (sub-records)
def _config_defs():
    """BASELINE.json configs[1..4] beside the headline one: shapes, comparisons, filter."""
    def c5_cats(S):
        rng = np.random.default_rng(5)
        cats = [[rng.integers(0, v, S).astype(np.int32), v] for v in (2, 5, 10, 20)]
        # the 2-value category follows the planted split (columns < S/2 vs the rest), so that planted rows survive
        # ratioFilter 0.8 and the sharded run has something to merge
        cats[0][0] = (np.arange(S) >= S // 2).astype(np.int32)
        return cats
    return {
        "C2b": dict(S=5_000, rows=1_000_000, snp=False, rf=0.0, sharded=False,
                    explicit=(np.arange(0, 300), np.arange(300, 700)), cats=lambda S: [],
                    what="1M x 5k binary, one comparison 300 vs 400 genomes: every test on the hypergeometric (FET) branch"),
        "C3": dict(S=10_000, rows=500_000, snp=True, rf=0.0, sharded=True, explicit=None,
                   cats=lambda S: [[(np.arange(S) * 4 // S).astype(np.int32), 4]],
                   what="SNP mode 500k sites x 10k genomes (2M allele rows), one 4-value category: 10 comparisons, dense output"),
        "C4": dict(S=20_000, rows=2_000_000, snp=False, rf=0.5, sharded=False, explicit=None,
                   cats=lambda S: [[(np.arange(S) // 400).astype(np.int32), 50]],
                   what="2M x 20k binary, one 50-value category: 1,275 comparisons (tcgen05 GEMM counts), ratioFilter 0.5"),
        "C5": dict(S=50_000, rows=10_000_000, snp=False, rf=0.8, sharded=True, explicit=None, cats=c5_cats,
                   what="scale run 10M x 50k binary, categories of 2/5/10/20 values: 281 comparisons, ratioFilter 0.8, "
                        "rows sharded over the GPUs (strong scaling), merge inside the timed region"),
    }


def run_config(name, cfg, capi, sharding, torch, dist, world, rank, local_rank, stream, barrier, steps, warmup, peaks,
               ncu):
    """One BASELINE config, device-resident, the whole job sharded by rows over the ranks when cfg['sharded'] (else
    rank 0 alone runs it at N=1 only).  Timed with CUDA events on the launching stream, max over ranks; for N>1 the
    gather + device merge of every comparison (K5b) is inside the timed region."""
    from helpers import make_spec
    S, rows_total, snp, rf = cfg["S"], cfg["rows"], cfg["snp"], cfg["rf"]
    dev = torch.device("cuda", local_rank)
    lo, hi = sharding.shard_bounds(rows_total, world)[rank]
    mult = 4 if snp else 1
    ctx = capi.Context(local_rank)
    try:
        ctx.set_stream(stream.cuda_stream)
        ctx.set_samples(S)
        ctx.set_row_base(lo * mult)
        ctx.reserve_rows(hi - lo)
        ctx.generate_synthetic(make_spec(capi.SynthSpec, S, law="uniform_half" if snp else "beta"), snp, lo, hi - lo)
        if cfg["explicit"] is not None:
            ctx.add_comparison(*cfg["explicit"])
        for vos, nv in cfg["cats"](S):
            ctx.add_category(vos, nv)
        n_comps = ctx.num_comparisons
        total_marker_rows = rows_total * mult
        merged_total = [0]

        merge_times = {}

        def step(times=None):
            ctx.run(capi.CORRECTION_BONFERRONI, rf, capi.SORT_ABS_RATIO_DESC, total_marker_rows)
            if world > 1:
                cols, seg = sharding.gather_and_merge_all_device(ctx, capi.SORT_ABS_RATIO_DESC, 0, want_device=True,
                                                                 times=times)
                if rank == 0:
                    merged_total[0] = int(seg[-1])
            else:
                merged_total[0] = ctx.result_total()

        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step()
            barrier()
            ctx.timings_accumulated(reset=True)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(steps):
                step()
            ev1.record(stream)
            barrier()
        acc = ctx.timings_accumulated(reset=True)
        t = ctx.timings()
        if world > 1:
            step(merge_times)     # one more, untimed step with synchronising laps: where the exchange spends its time
            ctx.timings_accumulated(reset=True)
        ms = torch.tensor([ev0.elapsed_time(ev1) / steps, acc["count_ms"] / steps, acc["test_ms"] / steps,
                           acc["order_ms"] / steps, acc["total_ms"] / steps], device=dev, dtype=torch.float64)
        surv = torch.tensor([ctx.result_total()], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(surv)
        ms = [float(x) for x in ms.tolist()]
        tests = total_marker_rows * n_comps
        hbm, bf16_burst = peaks
        rec = {"workload": cfg["what"], "marker_rows": total_marker_rows, "genomes": S, "comparisons": n_comps,
               "ratio_filter": rf, "n_gpus": world, "scaling": "strong" if world > 1 else None,
               "survivors": int(surv.item()), "merged_rows_rank0": merged_total[0] if world > 1 else None,
               "ms_per_step": ms[0], "tests_per_s": tests / (ms[0] * 1e-3),
               "stage_ms_max_over_ranks": {"count": ms[1], "test": ms[2], "order": ms[3], "library_total": ms[4],
                                           "merge_and_host": max(0.0, ms[0] - ms[4])},
               "engine": {1: "popc (K1)", 2: "gemm i8 (K2)", 3: "gemm mxf4 (K2)", 4: "gemm e5m2 (K2)"}.get(int(t["engine"]), "?"),
               "exchange_breakdown_rank0_ms": merge_times if world > 1 else None,
               "timer": "CUDA events on the launching stream around the steps (run + gather + merge), max over ranks"}
        # ---- rooflines of the stages, per rank (this rank's shard: bytes of its planes / its stage time)
        planes = 3 if snp else 2
        local_units = hi - lo
        count_bytes = local_units * ((S + 63) // 64) * 8 * planes
        local_count_ms = acc["count_ms"] / steps
        roofs = []
        gbs = count_bytes / (local_count_ms * 1e-3) / 1e9
        roofs.append({"kernel": "count stage (" + rec["engine"] + ")", "bound": "hbm", "achieved": gbs, "peak": hbm,
                      "unit": "GB/s", "frac": gbs / hbm, "algorithmic_bytes": count_bytes, "ms": local_count_ms})
        if int(t["engine"]) in (2, 3, 4):
            npad = 0
            for vos, nv in cfg["cats"](S):
                npad += 2 * ((nv + 1) // 2) + (2 if nv > 2 else 0)
            passes = -(-npad // 64)
            npad_tot = sum(-(-min(64, npad - 64 * q) // 16) * 16 for q in range(passes))
            ops = 2.0 * local_units * (-(-S // 128) * 128) * npad_tot * 2
            pops = ops / (local_count_ms * 1e-3) / 1e15
            fp4 = int(t["engine"]) == 3
            tpeak = 9.0 if fp4 else 4.5
            roofs.append({"kernel": "K2 tcgen05 GEMM counts (" + ("kind::mxf4" if fp4 else "kind::i8 / f8f6f4") + ")",
                          "bound": "tensor", "achieved": pops, "peak": tpeak,
                          "unit": "POP/s (MACs x 2 of both planes, N padded to 16)", "frac": pops / tpeak,
                          "peak_source": "nominal dense " + ("fp4 9" if fp4 else "int8 / fp8 4.5") + " POP/s (no measured figure for this type)",
                          "frac_of_int8_nominal_4.5": pops / 4.5,
                          "frac_of_2x_measured_bf16": pops / (2 * bf16_burst / 1e3),
                          "tensor_pipe_pct_ncu": (ncu.get("k2" if fp4 else "k2_i8") or {}).get("tensor_pipe_pct")})
        n_out = ctx.result_total()
        if n_out > 0 and acc["order_ms"] > 0:
            local_order_ms = acc["order_ms"] / steps
            passes = 7
            ob = n_out * 24.0 * passes
            ogbs = ob / (local_order_ms * 1e-3) / 1e9
            roofs.append({"kernel": "K4 radix order", "bound": "hbm", "achieved": ogbs, "peak": hbm, "unit": "GB/s",
                          "frac": ogbs / hbm, "items": n_out, "passes_assumed": passes,
                          "algorithmic_bytes": ob, "ms": local_order_ms,
                          "note": "12 B (key + index) read + 12 B written per item and 8-bit pass"})
        if name == "C2b" and world == 1:
            got = ctx.fetch(0)
            m = got["goa"].astype(np.int64) + got["gta"]
            k = got["goa"].astype(np.int64) + got["gob"]
            l = m + got["gob"] + got["gtb"]
            ok = (l > 0) & (k > 0) & ((got["gta"] + got["gtb"]) > 0) & (l < 1000)
            terms = np.where(ok, np.minimum(m, k) - np.maximum(0, m + k - l) + 1, 0).sum()
            flops = 4.0 * float(terms)
            local_test_ms = acc["test_ms"] / steps
            tf = flops / (local_test_ms * 1e-3) / 1e12
            roofs.append({"kernel": "K3 (screen + FET p-values)", "bound": "fp64", "achieved": tf, "peak": 40.0,
                          "unit": "TFLOP/s (4 fp64 ops per support term)", "frac": tf / 40.0,
                          "peak_source": "nominal FP64 vector ~40 TFLOP/s", "support_terms": int(terms), "ms": local_test_ms,
                          "fp64_pipe_pct_ncu": (ncu.get("k3b") or {}).get("fp64_pipe_pct")})
        rec["rooflines_rank0"] = roofs
        launches = int(acc["launches"])
        return rec, launches
    finally:
        ctx.close()


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
    ap.add_argument("--cpu-sample-rows", type=int, default=200_000)
    ap.add_argument("--ref-rows", type=int, default=ROWS_PER_GPU, help="--impl reference: marker rows per step")
    ap.add_argument("--configs", default="all", help="comma list of C2b,C3,C4,C5 sub-records, 'all' or 'none'")
    ap.add_argument("--config-steps", type=int, default=5)
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        # synchronise FIRST: dist.barrier() launches its NCCL kernel behind torch's current stream only, not behind the
        # stream the library runs on; a rank that had finished enqueuing would start the barrier kernel while its last
        # steps were still queued, and that kernel (spinning until every rank joins) keeps the cooperative sort from
        # getting all its CTAs resident -- the ranks then wait for one another (seen once at N = 8: +0.6 ms per window)
        torch.cuda.synchronize()
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
    launches_total = int(stage["launches"])
    ncu = _ncu_facts()

    # ------------------------------------------------------------------ end to end: `e2e`
    e2e = None
    parity = None
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
                merged, mseg = sharding.gather_and_merge_all_device(ctx, capi.SORT_ABS_RATIO_DESC, 0, out)
                assert merged is None or (len(merged["row"]) == total_rows and int(mseg[-1]) == total_rows)
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
        # ---- the rows the last e2e step delivered (all N GPUs' shards, merged) against the oracle
        if rank == 0:
            parity = parity_check_c2({k: v.numpy() for k, v in out.items()}, total_rows)
        host_rows = ctx.last_upload_host_rows_chars
        # bytes that crossed the link: raw chars for the DMA + K0 share, two packed planes for the host share
        h2d = (R - host_rows) * S + host_rows * ctx.row_stride_words * 16 + (R + 1) * 8 + (len(G1) + len(G2)) * 4
        # what the host cores alone need for their side of the upload (all ranks at once, no GPU involved): the
        # box's DRAM feeds N x 5 GB of 1-byte calls per step, which bounds e2e at N > 1
        Wd = ctx.row_stride_words
        sink = [np.empty((min(R, 200_000), Wd), dtype=np.uint64) for _ in range(2)]
        barrier()
        tp0 = time.perf_counter()
        for r0 in range(0, R, 200_000):
            n = min(200_000, R - r0)
            capi.lib().feht_pack_chars_host(chars.data_ptr() + r0 * S, offsets.data_ptr(), n, S, 0, Wd,
                                            sink[0].ctypes.data, sink[1].ctypes.data, None, host_threads + 1)
        host_pack_s = torch.tensor([time.perf_counter() - tp0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(host_pack_s, op=dist.ReduceOp.MAX)
        del sink
        e2e = {"value": total_rows / float(e2e_s.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "host_input_bytes_per_step": R * S,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * float(e2e_s.item()),
               "rank0_breakdown_ms": parts,
               "host_transcoder": {"threads": host_threads, "rows_packed_on_host": ctx.last_upload_host_rows_chars,
                                   "note": "feht_add_rows_chars splits the upload: DMA + GPU pack kernel for one end "
                                           "of the row range, AVX-512 / AVX2 host threads packing to 2 bits/call for the other"},
               "input_gbs_per_gpu": R * S / (parts["upload_pack_ms"] * 1e-3) / 1e9,
               "host_dram_ceiling": {"host_pack_only_ms": 1e3 * float(host_pack_s.item()),
                                     "host_read_gbs_all_ranks": world * R * S / float(host_pack_s.item()) / 1e9,
                                     "note": "all ranks transcoding their 5 GB of chars on their host cores at the same time, "
                                             "no GPU work: the rate at which this box's DRAM delivers the input"},
               "path": "feht_add_rows_chars(pinned host chars, 1 B/call; DMA+K0 and host transcoder threads) -> feht_run -> "
                       + ("feht_result_fetch" if world == 1 else
                          "feht_result_fetch_device(all comparisons), NCCL gather to rank 0, feht_merge_shards_device (K5b), one D2H"),
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
        del chars, out
    ctx.close()
    del ctx

    # ------------------------------------------------------------------ cpu baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        reps = 4     # ~16-20 core-seconds of CPU work in all
        v_all, dt_all = cpu_baseline(args.cpu_sample_rows, cores, reps=reps)
        one_rows = max(2000, args.cpu_sample_rows // 4)
        v_one, dt_one = cpu_baseline(one_rows, 1, reps=1)
        cpu = {"value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
               "single_thread_value": v_one,
               "sample": f"{args.cpu_sample_rows} of the 1,000,000 marker rows (same generator, same comparison), {reps} passes, "
                         f"oracle C port of the reference algorithm row-sharded over {cores} threads "
                         f"({dt_all:.2f} s wall = {dt_all * cores:.0f} core-seconds); single-thread figure on "
                         f"{one_rows} rows ({dt_one:.2f} s) -- the real feht binary is single-threaded and could not "
                         f"be built (no ghc/stack in the image); `--impl reference` runs all 1,000,000 rows"}

    # ------------------------------------------------------------------ the other BASELINE configs
    configs = {}
    want = [] if args.configs == "none" else (["C2b", "C3", "C4", "C5"] if args.configs == "all" else args.configs.split(","))
    torch.cuda.empty_cache()
    bf16 = 1644.9
    try:
        bf16 = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["bf16_tflops"])
    except Exception:
        pass
    for name, cfg in _config_defs().items():
        if name not in want or (world > 1 and not cfg["sharded"]):
            continue
        try:
            rec, n_l = run_config(name, cfg, capi, sharding, torch, dist, world, rank, local_rank, stream, barrier,
                                  args.config_steps, 3, (peak, bf16), ncu)
            configs[name] = rec
        except capi.FehtError as e:      # reported, not hidden: the headline line still prints
            configs[name] = {"error": str(e)}
        torch.cuda.empty_cache()

    rc = 0
    if rank == 0:
        k1 = ncu.get("k1") or {}
        traffic = None
        if R == ROWS_PER_GPU and "dram_bytes_read" in k1:
            traffic = float(k1["dram_bytes_read"]) + float(k1["dram_bytes_write"])
        elif R == ROWS_PER_GPU:
            traffic = _k1_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 popcount + f64", "data": "synthetic",
            "config": _config(world, R),
            "stage_ms_per_step": {k: stage[k] / args.steps for k in ("count_ms", "test_ms", "order_ms", "total_ms")},
            "timed_window": "CUDA events on the stream the runs are enqueued on; ms_per_step >= sum of the library's own "
                            "stage events (asserted)",
            "roofline": {"bound": "hbm", "kernel": "count_popc_kernel (K1)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "traffic_source": k1.get("source", "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per "
                                                            "launch (profiles/k1_traffic.json)"),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": count_bytes, "avg_launch_ms": count_ms},
            "cpu_baseline": cpu, "e2e": e2e, "parity_check": parity,
            "configs": configs,
            "gpu_launches": launches_total,
            "clocks": clocks,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
        if parity is not None and parity["mismatches"]:
            sys.stderr.write(f"PARITY CHECK FAILED: {parity}\n")
            rc = 1
    if world > 1:
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
