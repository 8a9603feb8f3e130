"""One tcgen05 conv forward + backward of a refine layer (128->128 3x3, N x HW x HW) for `ncu --set full`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import isr_b200  # noqa: F401,E402
from isr_b200 import training as T  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
x = torch.randn(N, 128, HW, HW, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_()
w = (torch.randn(128, 128, 3, 3, device=dev) / 34.0).requires_grad_()
b = torch.zeros(128, device=dev, requires_grad=True)
gy = torch.randn(N, 128, HW, HW, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
for _ in range(2):
    y = T.conv2d(x, w, b, tc=True, out_bf16=True)
    y.backward(gy)
torch.cuda.synchronize()
print("ok")
