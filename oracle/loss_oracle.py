"""CPU oracle for the stage-2/3 training losses (TEST INFRASTRUCTURE ONLY).

Plain torch CPU restatement of the reference's ``L1Loss``, ``SSIMLoss``, ``FFTLoss`` and the
``SWTLoss`` GPU-approximation path (``src/losses/perceptual_loss.py``), differentiable by
autograd so the tests can compare both loss values and d(loss)/d(pred) with the fused CUDA
kernels.  Only ``tests/`` and ``bench.py``'s CPU-baseline legs may import this module.

Parity pin: ``tests/golden/losses_24x24.npz`` holds the four loss values the *reference
classes themselves* produce on a fixed (pred, target) pair (``oracle/make_golden.py``);
``tests/test_oracle_golden.py`` checks this restatement against them.
All citations are relative to ``/root/reference``.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

# configs/train_config.yaml:140-175
STAGE_WEIGHTS = {
    1: {"l1": 1.0},
    2: {"l1": 0.75, "swt": 0.20, "ssim": 0.05},
    3: {"l1": 0.60, "swt": 0.25, "fft": 0.10, "ssim": 0.05},
}


def l1_loss(pred, target):
    """L1Loss.forward, perceptual_loss.py:86-104."""
    return (pred - target).abs().mean()


def _gauss_window(size=11, sigma=1.5, channels=3):
    """SSIMLoss._create_gaussian_window, perceptual_loss.py:225-241."""
    g = torch.tensor([math.exp(-(x - size // 2) ** 2 / (2 * sigma ** 2)) for x in range(size)], dtype=torch.float32)
    g = g / g.sum()
    w2 = g[:, None].mm(g[None, :])[None, None]
    return w2.expand(channels, 1, size, size).contiguous()


def ssim_loss(pred, target):
    """SSIMLoss._ssim/forward, perceptual_loss.py:243-291 (zero pad 5, C1=1e-4, C2=9e-4)."""
    Cc = pred.shape[1]
    w = _gauss_window(11, 1.5, Cc).to(device=pred.device, dtype=pred.dtype)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mu1 = F.conv2d(pred, w, padding=5, groups=Cc)
    mu2 = F.conv2d(target, w, padding=5, groups=Cc)
    s1 = F.conv2d(pred * pred, w, padding=5, groups=Cc) - mu1 ** 2
    s2 = F.conv2d(target * target, w, padding=5, groups=Cc) - mu2 ** 2
    s12 = F.conv2d(pred * target, w, padding=5, groups=Cc) - mu1 * mu2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 ** 2 + mu2 ** 2 + C1) * (s1 + s2 + C2))
    return 1 - m.mean()


def fft_loss(pred, target, high_freq_weight=2.0):
    """FFTLoss.forward with focus_high_freq=True, perceptual_loss.py:533-598."""
    H, W = pred.shape[-2:]
    P = torch.fft.fftshift(torch.fft.fft2(pred, norm="ortho"), dim=(-2, -1))
    T = torch.fft.fftshift(torch.fft.fft2(target, norm="ortho"), dim=(-2, -1))
    cy, cx = H // 2, W // 2
    yy, xx = torch.meshgrid(torch.arange(H).float() - cy, torch.arange(W).float() - cx, indexing="ij")
    wts = (1.0 + (high_freq_weight - 1.0) * torch.sqrt(xx ** 2 + yy ** 2) / math.sqrt(cy ** 2 + cx ** 2)).to(device=pred.device, dtype=pred.dtype)
    mag = (P.abs() - T.abs()).abs() * wts
    ph = (P.angle() - T.angle()).abs() * wts
    return mag.mean() + 0.1 * ph.mean()


def _haar_filters(dtype):
    """SWTLoss._init_wavelet_filters, perceptual_loss.py:661-682 with pywt's haar taps
    dec_lo = [s, s], dec_hi = [-s, s], s = 1/sqrt(2)."""
    s = 1.0 / math.sqrt(2.0)
    lo = torch.tensor([s, s], dtype=torch.float32)
    hi = torch.tensor([-s, s], dtype=torch.float32)
    f = torch.stack([lo[None] * lo[:, None], lo[None] * hi[:, None], hi[None] * lo[:, None], hi[None] * hi[:, None]])
    return f[:, None].to(dtype)


def swt_coeffs(x, level=2):
    """SWTLoss._swt2d_gpu, perceptual_loss.py:684-733."""
    B, Cc, H, W = x.shape
    filt = _haar_filters(x.dtype).to(x.device)
    out, cur = [], x
    for lv in range(level):
        pad = 2 ** lv
        p = F.pad(cur, (pad, pad, pad, pad), mode="reflect")
        c = F.conv2d(p.reshape(B * Cc, 1, *p.shape[-2:]), filt, dilation=2 ** lv)
        c = c.reshape(B, Cc, 4, *c.shape[-2:])[..., :H, :W]
        cA, cH, cV, cD = c[:, :, 0], c[:, :, 1], c[:, :, 2], c[:, :, 3]
        out.append((cA, cH, cV, cD))
        cur = cA
    return out


def swt_loss(pred, target, level=2):
    """SWTLoss._forward_gpu, perceptual_loss.py:797-813 (band weights a .5 / h 1.5 / v 1.5 / d 2)."""
    bw = (0.5, 1.5, 1.5, 2.0)
    pc, tc = swt_coeffs(pred, level), swt_coeffs(target, level)
    loss = 0.0
    for lv in range(level):
        for k in range(4):
            loss = loss + bw[k] * (pc[lv][k] - tc[lv][k]).abs().mean()
    return loss / level


LOSSES = {"l1": l1_loss, "ssim": ssim_loss, "fft": fft_loss, "swt": swt_loss}


def combined_loss(pred, target, weights: Dict[str, float]) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """CombinedLoss.forward restricted to the four stage-1/2/3 components, weight-driven
    (perceptual_loss.py:1207-1284): a component is evaluated iff its weight > 0."""
    comps, total = {}, 0.0
    for name in ("l1", "ssim", "fft", "swt"):
        w = weights.get(name, 0.0)
        if w > 0:
            comps[name] = LOSSES[name](pred, target)
            total = total + w * comps[name]
    return total, comps


def fft_branch_cut_slack(pred, target, high_freq_weight=2.0) -> float:
    """Upper bound on how far two correct evaluations of ``fft_loss`` may differ.

    On the self-conjugate bins (k == -k mod n: DC / Nyquist rows and columns) the spectrum of a real
    image is real up to rounding noise.  Where it is NEGATIVE the reference's ``torch.angle`` returns
    +pi or -pi depending on the sign of that noise, so ``|angle P - angle T|`` is 0 or 2*pi at random
    (the reference's own CPU and CUDA builds disagree there).  Each such bin can move the loss by
    0.1 * w * 2*pi / n; this returns the sum over the affected bins (fp64 spectrum)."""
    B, Cc, H, W = pred.shape
    P = torch.fft.fft2(pred.double(), norm="ortho")
    T = torch.fft.fft2(target.double(), norm="ortho")
    ys = [0] + ([H // 2] if H % 2 == 0 else [])
    xs = [0] + ([W // 2] if W % 2 == 0 else [])
    cy, cx = H // 2, W // 2
    slack = 0.0
    for ky in ys:
        for kx in xs:
            sy, sx = (ky + cy) % H - cy, (kx + cx) % W - cx
            w = 1.0 + (high_freq_weight - 1.0) * math.sqrt(sx * sx + sy * sy) / math.sqrt(cy * cy + cx * cx)
            neg = (P[..., ky, kx].real < 0) | (T[..., ky, kx].real < 0)
            slack += float(neg.sum()) * 0.1 * w * 2 * math.pi
    return slack / pred.numel()


# ----------------------------------------------------------------------------------------------
# validation metrics (src/utils/metrics.py) -- oracle for image-super-resolution_b200/metrics.py
# ----------------------------------------------------------------------------------------------
def metric_rgb_to_y(img):
    """metrics.py:30-52."""
    r, g, b = img[:, 0:1], img[:, 1:2], img[:, 2:3]
    return (65.481 * r + 128.553 * g + 24.966 * b + 16.0) / 255.0


def _metric_prep(a, b, crop_border, test_y_channel):
    a, b = a.clamp(0, 1), b.clamp(0, 1)
    if a.dim() == 3:
        a, b = a[None], b[None]
    if crop_border > 0:
        a = a[:, :, crop_border:-crop_border, crop_border:-crop_border]
        b = b[:, :, crop_border:-crop_border, crop_border:-crop_border]
    if test_y_channel and a.shape[1] == 3:
        a, b = metric_rgb_to_y(a), metric_rgb_to_y(b)
    return a, b


def metric_psnr(a, b, crop_border=0, test_y_channel=False) -> float:
    """calculate_psnr, metrics.py:75-126."""
    a, b = _metric_prep(a, b, crop_border, test_y_channel)
    mse = float(((a - b) ** 2).mean())
    return float("inf") if mse < 1e-10 else 10 * math.log10(1.0 / mse)


def metric_ssim(a, b, crop_border=0, test_y_channel=False) -> float:
    """calculate_ssim on its torch path (calculate_ssim_torch, metrics.py:128-186): the SSIM-loss map mean."""
    a, b = _metric_prep(a, b, crop_border, test_y_channel)
    return float(1 - ssim_loss(a, b))
