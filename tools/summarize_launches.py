"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/summarize_launches.py profiles/r01_fp32_launches.csv [skip_first_n_launches]
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))[skip:]
tot = collections.OrderedDict()
for x in rows:
    name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    v = float(x["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e6, "us": v / 1e3, "usecond": v / 1e3, "nsecond": v / 1e6}.get(x["Metric Unit"], v)
    t = tot.setdefault(name, [0.0, 0])
    t[0] += v
    t[1] += 1
s = sum(v[0] for v in tot.values())
print(f"{len(rows)} launches, {s:.2f} ms total (per-launch times are cold-cache and serialised: compare shares)")
print(f"{'ms':>10} {'share':>6} {'n':>4}  kernel")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]:10.3f} {100 * v[0] / s:5.1f}% {v[1]:4d}  {k[:100]}")
