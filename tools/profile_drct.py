"""One warm-up + one DRCT-L forward at 352x512 (the command ncu wraps).   python tools/profile_drct.py [fp32|bf16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr_b200 import drct as D
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = D.create_drct_model().to(dev).eval()
m.precision = prec
x = torch.rand(1, 3, 352, 512, generator=torch.Generator().manual_seed(1234)).to(dev)
with torch.no_grad():
    m(x)
    torch.cuda.synchronize()
    y = m(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
