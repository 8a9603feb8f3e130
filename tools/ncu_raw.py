"""Print selected raw metrics of an .ncu-rep (first kernel): python tools/ncu_raw.py file.ncu-rep [substr ...]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
keys = sys.argv[2:] or ["issue_stalled", "gpu__time_duration", "inst_executed.sum", "pipe_tensor", "cycles_active.avg", "warps_active"]
for h, u, v in zip(rows[0], rows[1], rows[2]):
    if any(k in h for k in keys):
        try:
            if float(v.replace(",", "")) == 0:
                continue
        except ValueError:
            pass
        print(f"{h:110s} {u:14s} {v}")
