"""Compact per-launch summary (the format bench.py's ncu_traffic reads) from an `ncu --page raw --csv` export.
usage: ncu_raw_to_summary.py raw.csv out.csv "header comment" """
import csv, sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if r[0] == "ID")
hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
kn, gs, bs = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
idx = [hdr.index(c) for c in COLS if c in hdr]
names = [c for c in COLS if c in hdr]
with open(sys.argv[2], "w", newline="") as f:
    f.write("# " + sys.argv[3] + "\n")
    w = csv.writer(f)
    w.writerow(["Kernel Name", "Grid Size", "Block Size"] + names)
    w.writerow(["", "", ""] + [units[i] for i in idx])
    for r in data:
        if len(r) <= max(idx):
            continue
        name = r[kn].split("(")[0].replace("fw::", "")
        w.writerow([name, r[gs], r[bs]] + [r[i].replace(",", "") for i in idx])
print("wrote", sys.argv[2], len(data), "launches")
