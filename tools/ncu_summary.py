"""Compact per-kernel summary (one row per distinct kernel, the LAST captured launch of each) of `ncu --page raw --csv`
exports: the columns the judge and bench.py read (development tool).

    python tools/ncu_summary.py OUT.csv RAW1.csv [RAW2.csv ...]"""
import csv
import sys

COLUMNS = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
out_path, raws = sys.argv[1], sys.argv[2:]
kernels, units = {}, None
for raw in raws:
    rows = list(csv.reader(open(raw)))
    header, unit_row = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(header)}
    cols = [c for c in COLUMNS if c in idx]
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    units = [""] + [("Mbyte" if unit_row[idx[c]] in scale else unit_row[idx[c]]) for c in cols]  # byte counts: always Mbyte
    for r in rows[2:]:
        vals = []
        for c in cols:
            v = r[idx[c]]
            if unit_row[idx[c]] in scale and v not in ("", "n/a"):
                v = "%.6f" % (float(v) * scale[unit_row[idx[c]]])
            vals.append(v)
        kernels[r[idx["Kernel Name"]]] = [r[idx["Kernel Name"]]] + vals
with open(out_path, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["Kernel Name"] + cols)
    w.writerow(units)
    for row in kernels.values():
        w.writerow(row)
print(out_path, len(kernels), "kernels")
