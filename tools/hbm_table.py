"""Per-kernel HBM roofline table from an ncu launch list that carries gpu__time_duration.sum,
dram__bytes_read.sum and dram__bytes_write.sum (one row per metric per launch):

    python tools/hbm_table.py launches.csv [anchor kernel substring] [peak GB/s]

Only the launches from the LAST occurrence of the anchor kernel on are counted (one forward / one step).
achieved = measured DRAM bytes / duration; the launch list is cold-cache and serialised, so these are lower
bounds on what the kernels reach inside a step.
"""
import collections
import csv
import json
import os
import re
import sys

path = sys.argv[1]
anchor = sys.argv[2] if len(sys.argv) > 2 else "k_dct_bands"
peak = float(sys.argv[3]) if len(sys.argv) > 3 else None
if peak is None:
    try:
        mp = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        peak = float(mp.get("hbm_gbps_sustained") or mp.get("hbm_gbps") or 6555.0)
    except Exception:
        peak = 6555.0
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"].startswith("gpu__time"):
        d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(u, 1e-6)
    else:
        b = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d["bytes"] = d.get("bytes", 0.0) + b
ids = list(launch)
last = max((i for i, k in enumerate(ids) if anchor in launch[k]["name"]), default=0)
agg = collections.OrderedDict()
for k in ids[last:]:
    d = launch[k]
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, [0.0, 0.0, 0])
    a[0] += d.get("ms", 0.0)
    a[1] += d.get("bytes", 0.0)
    a[2] += 1
tot = sum(a[0] for a in agg.values())
print(f"{len(ids) - last} launches from the last '{anchor}', {tot:.2f} ms; HBM peak used for %: {peak:.0f} GB/s")
print(f"{'ms':>8} {'share':>6} {'n':>4} {'DRAM MB':>9} {'GB/s':>8} {'%peak':>6}  kernel")
for name, (ms, b, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    gbs = b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    print(f"{ms:8.3f} {100 * ms / tot:5.1f}% {n:4d} {b / 1e6:9.1f} {gbs:8.0f} {100 * gbs / peak:5.1f}%  {name[:90]}")
