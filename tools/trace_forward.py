"""In-situ duration of every launch of a fusion forward: CUDA events around each library call inside back-to-back
forwards (single stream), so producer outputs are still in the 126 MB L2 when the consumer runs -- unlike an ncu launch
list, which is cold-cache and serialised.  Event pairs add a few microseconds of gap per launch.
    python tools/trace_forward.py [--lr 339 510] [--iters 5] [--precision bf16] [--min-ms 0.05]
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import isr_b200  # noqa: E402
from oracle import fusion_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lr", type=int, nargs=2, default=[339, 510])
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--min-ms", type=float, default=0.0)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = isr_b200.CompleteEnhancedFusionSR(None).eval().to(dev)
m.precision = a.precision
lr, imgs, fts, _ = O.synthetic_inputs(1, a.lr[0], a.lr[1])
lr, imgs, fts = lr.to(dev), {k: v.to(dev) for k, v in imgs.items()}, {k: v.to(dev) for k, v in fts.items()}
for _ in range(3):
    m.forward_with_precomputed(lr, imgs, fts)
eng = m._engine
eng.overlap_routing = False
torch.cuda.synchronize()
runs = []
for _ in range(a.iters):
    eng.trace = []
    m.forward_with_precomputed(lr, imgs, fts)
    runs.append(eng.trace)
eng.trace = None
torch.cuda.synchronize()
n = len(runs[0])
tot = 0.0
print(f"{'ms':>8}  launch (mean of {a.iters} forwards, in launch order)")
for i in range(n):
    label = runs[0][i][0]
    ms = sum(r[i][1].elapsed_time(r[i][2]) for r in runs) / a.iters
    tot += ms
    if ms >= a.min_ms:
        print(f"{ms:8.3f}  {label}")
span = sum(r[0][1].elapsed_time(r[-1][2]) for r in runs) / a.iters
print(f"sum of launches {tot:.3f} ms; first-to-last span {span:.3f} ms; {n} traced calls")
