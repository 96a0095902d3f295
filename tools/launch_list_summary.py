"""Turn the CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv` into the per-kernel share table
kept under profiles/: python tools/launch_list_summary.py gpurun_out/launches.csv > profiles/rNN_ncu_launch_list.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
tot, cnt, order = collections.OrderedDict(), collections.Counter(), []
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("rtdd::", "")
    us = float(r[vi].replace(",", "")) / 1000.0
    tot[name] = tot.get(name, 0.0) + us
    cnt[name] += 1
    order.append((name, r[gi], r[bi], us))
T = sum(tot.values())
print("# total %.1f us over %d launches; per-launch times are cold-cache and serialised: compare SHARES" % (T, len(order)))
print()
print("%-52s %6s %12s %7s" % ("kernel", "count", "total us", "share"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-52s %6d %12.1f %6.1f%%" % (k, cnt[k], v, 100 * v / T))
print()
print("# in launch order (kernel, grid, block, us)")
for name, g, b, us in order:
    print("%-52s %-18s %-18s %8.1f" % (name, g, b, us))
