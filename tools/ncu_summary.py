"""Summarise an `ncu --set full` report: per kernel, the median captured launch for the metrics that matter
here. Usage: python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv, io, re, statistics as st, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_atom.sum",
    "lts__t_sectors_op_red.sum", "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
kcol = col["Kernel Name"]
by = {}
for r in data:
    name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[kcol]).split("::")[-1])
    by.setdefault(name, []).append(r)
for name, rs in by.items():
    print(f"=== {name} captured launches = {len(rs)}")
    for m in METRICS:
        if m not in col:
            continue
        vals = []
        for r in rs:
            try:
                vals.append(float(r[col[m]].replace(",", "")))
            except ValueError:
                pass
        if vals:
            print(f"   {m:80s} {st.median(vals):.6f} {units[col[m]]}")
