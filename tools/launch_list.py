"""Per-launch durations of the LAST forward in an ncu gpu__time_duration launch list.
    python tools/launch_list.py gpurun_out/launches.csv [min_ms]
"""
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
idx = [i for i, r in enumerate(rows) if "k_dct_bands" in r["Kernel Name"]]
sec = rows[idx[-1]:]
tot = 0.0
for r in sec:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r["Metric Value"].replace(",", "")) / 1e6
    tot += v
    if v >= thr:
        print(f"{v:8.3f} ms  grid={r['Grid Size']:>14s}  {name[:70]}")
print(f"total {tot:.3f} ms over {len(sec)} launches")
