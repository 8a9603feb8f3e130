"""Micro-benchmark of the tcgen05 training conv kernels (forward, input gradient, weight gradient) on the
layer shapes of the training workloads.  CUDA-event timing, 3 warm-up + 10 timed launches per kernel.
    python tools/bench_conv_train.py [N] [HW]          (default N=32, HW=256: the C2 refine layers)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import isr_b200  # noqa: F401,E402
from isr_b200 import training as T  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 256
only = sys.argv[3] if len(sys.argv) > 3 else ""
dev = torch.device("cuda:0")
shapes = [(128, 128, 3), (76, 64, 3), (64, 64, 3), (64, 32, 3), (32, 32, 3), (96, 32, 3), (3, 128, 3), (128, 3, 3), (32, 16, 3)]
if only:
    ci, co, ks = (int(v) for v in only.split(","))
    shapes = [(ci, co, ks)]


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"N={N} HxW={HW}x{HW}")
for ci, co, ks in shapes:
    x = torch.randn(N, ci, HW, HW, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_()
    w = (torch.randn(co, ci, ks, ks, device=dev) / (ci * ks * ks) ** 0.5).requires_grad_()
    b = torch.zeros(co, device=dev, requires_grad=True)
    gy = torch.randn(N, co, HW, HW, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    flop = 2.0 * ks * ks * ci * co * N * HW * HW
    y = T.conv2d(x, w, b, tc=True, out_bf16=True)
    t_f = timed(lambda: T.conv2d(x, w, b, tc=True, out_bf16=True))
    # backward pieces separately
    t_d = timed(lambda: torch.autograd.grad(y, x, gy, retain_graph=True))
    t_w = timed(lambda: torch.autograd.grad(y, w, gy, retain_graph=True))
    print(f"{ci:4d}->{co:4d} k{ks}: fwd {t_f:7.3f} ms {flop / t_f / 1e9:7.1f} TF/s | dgrad {t_d:7.3f} ms {flop / t_d / 1e9:7.1f} TF/s"
          f" | wgrad {t_w:7.3f} ms {flop / t_w / 1e9:7.1f} TF/s")
