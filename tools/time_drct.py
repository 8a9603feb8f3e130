"""DRCT-L forward time at 352x512 in the bf16 mode (CUDA events, mean of N forwards) + max-abs against the fp32 mode.
    python tools/time_drct.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr_b200 import drct as D
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = D.create_drct_model().to(dev).eval()
x = torch.rand(1, 3, 352, 512, generator=torch.Generator().manual_seed(1234)).to(dev)
with torch.no_grad():
    m.precision = "bf16"
    for _ in range(2):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    m.precision = "fp32"
    ref = m(x)
    torch.cuda.synchronize()
print(f"DRCT-L bf16 forward {ms:.2f} ms per 352x512 image; max-abs vs fp32 mode {(y.float() - ref.float()).abs().max().item():.3e}")
