"""Where a sweep_cluster_kernel launch spends its time: warp-stall samples and executed instructions of an `ncu --set full
--import-source on` capture, bucketed by phase of the per-region loop.  The phase boundaries are found from marker instructions
in the SASS (tile mbarrier wait, the barrier that ends the prologue, the sweep barriers, the first store of the write-back).
    ncu -i capture.ncu-rep --page source --csv --print-source sass > sass.csv ; python tools/ncu_phase_breakdown.py sass.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
hdr = rows[starts[0] + 1]
end = starts[1] if len(starts) > 1 else len(rows)
data = [r for r in rows[starts[0] + 2:end] if len(r) > 10]
H = {h: i for i, h in enumerate(hdr)}
src = [r[1] for r in data]
smp = [int(r[H["# Samples"]]) for r in data]
ins = [int(r[H["Instructions Executed"]]) for r in data]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def first(pred, lo=0):
    for k in range(lo, len(src)):
        if pred(src[k]):
            return k
    return len(src)


per_cta = max(ins)                                   # the sweep loop's instructions run most often
tile_wait = first(lambda s: "SYNCS.PHASECHK" in s)
prologue = tile_wait + 13
bars = [k for k, s in enumerate(src) if "BAR." in s and "DEFER_BLOCKING" in s and ins[k] > 0 and k > prologue]
prologue_end = bars[0] - 30
sweeps = first(lambda s: "UTMALDG" in s, bars[0])
sweeps = [k for k, s in enumerate(src) if "UTMALDG" in s][-1] + 6
last_bar = [k for k in bars if ins[k] > 0][-1]
writeback = last_bar + 8
exit_ = first(lambda s: "EXIT" in s, writeback)
phases = [("set-up (once per CTA)", 0, tile_wait - 110), ("region head", tile_wait - 110, tile_wait), ("wait for the tile", tile_wait, prologue),
          ("prologue", prologue, prologue_end), ("edge rows, barrier, next tile issued", prologue_end, sweeps),
          ("sweeps", sweeps, writeback), ("write-back", writeback, exit_ - 3), ("exit", exit_ - 3, len(src))]
tot_s, tot_i = sum(smp), sum(ins)
print("kernel:", rows[starts[0]][1][:100])
print("warp-stall samples %d, warp instructions %d" % (tot_s, tot_i))
for name, a, b in phases:
    s, i = sum(smp[a:b]), sum(ins[a:b])
    st = {h: sum(int(r[H[h]]) for r in data[a:b]) for h in stalls}
    top = sorted(st.items(), key=lambda x: -x[1])[:4]
    print("%-40s %5.1f %% of samples  %5.1f %% of instructions   %s" % (name, 100.0 * s / tot_s, 100.0 * i / tot_i,
                                                                      " ".join("%s=%d" % (k[6:], v) for k, v in top)))
