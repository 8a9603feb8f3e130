"""One warm-up + N fusion forwards at a given LR size: the short command ncu wraps.
    python tools/profile_forward.py [--lr 339 510] [--iters 1] [--precision fp32|bf16]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import isr_b200  # noqa: E402
from oracle import fusion_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lr", type=int, nargs=2, default=[339, 510])
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--precision", default="fp32")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = isr_b200.CompleteEnhancedFusionSR(None).eval().to(dev)
m.precision = a.precision
lr, imgs, fts, _ = O.synthetic_inputs(1, a.lr[0], a.lr[1])
lr, imgs, fts = lr.to(dev), {k: v.to(dev) for k, v in imgs.items()}, {k: v.to(dev) for k, v in fts.items()}
m.forward_with_precomputed(lr, imgs, fts)          # warm-up (weight packing, workspaces)
torch.cuda.synchronize()
for _ in range(a.iters):
    sr = m.forward_with_precomputed(lr, imgs, fts)
torch.cuda.synchronize()
print("ok", tuple(sr.shape), m._engine.launches, "launches/forward")
