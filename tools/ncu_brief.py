"""Print the handful of ncu metrics the profiles/ notes quote, per kernel launch of a .ncu-rep:
python tools/ncu_brief.py report.ncu-rep"""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.per_cycle_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:90])
    for k in KEYS:
        if k in d: print("  %-80s %s %s" % (k, d[k], units[hdr.index(k)]))
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(d[h] or 0) > 0.15:
            print("  stall %-60s %s" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], d[h]))
