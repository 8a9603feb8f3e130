"""Runs N training steps (default 2) of a workload for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_train.py c2 8 fp32
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import isr_b200  # noqa: E402
from isr_b200.losses import CombinedLoss  # noqa: E402
from isr_b200.trainer import FusionTrainer  # noqa: E402
from oracle import fusion_oracle as O  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
hw = 64 if wl == "c2" else 96
w = {"l1": 1.0} if wl == "c2" else {"l1": 0.60, "swt": 0.25, "fft": 0.10, "ssim": 0.05}
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = isr_b200.CompleteEnhancedFusionSR(None).to(dev)
m.precision = prec
crit = CombinedLoss()
crit.set_weights({"charbonnier": 0, "l2": 0, "vgg": 0, "edge": 0, "clip": 0, "swt": 0, "fft": 0, "ssim": 0, **w})
tr = FusionTrainer(m, crit)
lr, imgs, fts, hr = O.synthetic_inputs(B, hw, hw)
args = (lr.to(dev), {k: v.to(dev) for k, v in imgs.items()}, {k: v.to(dev) for k, v in fts.items()}, hr.to(dev))
for i in range(steps):
    torch.cuda.synchronize()
    loss, _ = tr.step(*args)
torch.cuda.synchronize()
print("loss", float(loss))
