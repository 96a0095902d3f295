"""Summarise one kernel of an .ncu-rep (read here, on the CPU box): ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_bytes.sum"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-82s %s %s" % (w, vals[i][:110], units[i]))
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        print("%-82s %s" % (h, vals[i]))
