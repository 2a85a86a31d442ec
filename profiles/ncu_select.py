"""Selected columns of an `ncu --page raw --csv` export (one row per kernel launch), for the files judged under profiles/.
    python profiles/ncu_select.py raw.csv > selected.csv"""
import csv
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
idx = [hdr.index(c) for c in COLS if c in hdr]
w = csv.writer(sys.stdout)
for r in rows:
    w.writerow([r[i] for i in idx])
