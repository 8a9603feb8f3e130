"""Validation metrics on the device: Y-channel PSNR / SSIM with border crop, as
``src/utils/metrics.py:30-126, 128-186`` computes them, without the per-image ``.item()`` syncs and host round
trips of ``validate_epoch`` (train.py:450-515).  Results stay device tensors (one value per image); the
reference-named wrappers ``calculate_psnr`` / ``calculate_ssim`` return Python floats like the originals.

SSIM follows the reference's torch path (``calculate_ssim_torch``: 11x11 Gaussian sigma 1.5, zero padding, C1 = 1e-4,
C2 = 9e-4), which is what ``calculate_ssim`` uses when scikit-image is not installed; it is evaluated by the fused
SSIM kernel of ``losses.py`` (``ffsr_loss_ssim``).
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

from .losses import fused_losses


def rgb_to_y(img: torch.Tensor) -> torch.Tensor:
    """ITU-R BT.601 luminance in [16/255, 235/255] (metrics.py:30-52)."""
    r, g, b = img[..., 0:1, :, :], img[..., 1:2, :, :], img[..., 2:3, :, :]
    return (65.481 * r + 128.553 * g + 24.966 * b + 16.0) / 255.0


def _prep(a: torch.Tensor, b: torch.Tensor, crop_border: int, test_y_channel: bool):
    assert a.shape == b.shape, f"Image shapes must match: {a.shape} vs {b.shape}"
    a, b = a.clamp(0, 1), b.clamp(0, 1)
    if a.dim() == 3:
        a, b = a.unsqueeze(0), b.unsqueeze(0)
    if crop_border > 0:
        a = a[:, :, crop_border:-crop_border, crop_border:-crop_border]
        b = b[:, :, crop_border:-crop_border, crop_border:-crop_border]
    if test_y_channel and a.size(1) == 3:
        a, b = rgb_to_y(a), rgb_to_y(b)
    return a.float().contiguous(), b.float().contiguous()


@torch.no_grad()
def psnr_ssim_per_image(sr: torch.Tensor, hr: torch.Tensor, crop_border: int = 4, test_y_channel: bool = True
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(psnr[B] in dB, ssim[B]) as device tensors; nothing here synchronises the host."""
    if not sr.is_cuda:
        raise RuntimeError("metrics (sm_100a build) need CUDA tensors: there is no CPU path")
    a, b = _prep(sr, hr, crop_border, test_y_channel)
    mse = ((a - b) ** 2).mean(dim=(1, 2, 3))
    psnr = torch.where(mse < 1e-10, torch.full_like(mse, float("inf")), 10.0 * torch.log10(1.0 / mse.clamp_min(1e-10)))
    ssim = torch.stack([1.0 - fused_losses(a[i:i + 1], b[i:i + 1], {"ssim": 1.0})[0] for i in range(a.shape[0])])
    return psnr, ssim


def calculate_psnr(img1, img2, crop_border: int = 0, test_y_channel: bool = False) -> float:
    """metrics.py:75-126 (mean over the whole batch; inf below 1e-10 MSE)."""
    a, b = _prep(img1, img2, crop_border, test_y_channel)
    mse = float(((a - b) ** 2).mean())
    return float("inf") if mse < 1e-10 else 10.0 * math.log10(1.0 / mse)


def calculate_ssim(img1, img2, crop_border: int = 0, test_y_channel: bool = False) -> float:
    """metrics.py:189-246, torch path (mean of the SSIM map over batch, channels and pixels)."""
    a, b = _prep(img1, img2, crop_border, test_y_channel)
    with torch.no_grad():
        return float(1.0 - fused_losses(a, b, {"ssim": 1.0})[0])


def calculate_psnr_ssim_batch(sr_images, hr_images, crop_border: int = 4, test_y_channel: bool = True) -> Tuple[float, float]:
    """metrics.py:249-280: average PSNR / SSIM over the images of a batch (one sync at the end)."""
    p, s = psnr_ssim_per_image(sr_images, hr_images, crop_border, test_y_channel)
    finite = torch.isfinite(p)                                 # identical images are left out of the PSNR average
    avg_p = float(p[finite].mean()) if bool(finite.any()) else float("inf")
    return avg_p, float(s.mean())


class MetricCalculator:
    """``src.utils.metrics.MetricCalculator`` (metrics.py:291-375) without the two host syncs per image: ``update`` keeps
    the per-image PSNR / SSIM as device tensors, ``get_metrics`` reduces them with ONE synchronisation.  Same results:
    infinite PSNR values are left out of the PSNR mean, every image counts for SSIM, empty -> zeros."""

    def __init__(self, crop_border: int = 4, test_y_channel: bool = True):
        self.crop_border, self.test_y_channel = crop_border, test_y_channel
        self.reset()

    def reset(self) -> None:
        self._psnr, self._ssim, self.count = [], [], 0

    @torch.no_grad()
    def update(self, sr: torch.Tensor, hr: torch.Tensor) -> None:
        p, s = psnr_ssim_per_image(sr, hr, self.crop_border, self.test_y_channel)
        self._psnr.append(p)
        self._ssim.append(s)
        self.count += int(sr.shape[0]) if sr.dim() == 4 else 1

    def get_metrics(self):
        if self.count == 0:
            return {"psnr": 0.0, "ssim": 0.0}
        p, s = torch.cat(self._psnr), torch.cat(self._ssim)
        finite = torch.isfinite(p)
        both = torch.stack([torch.where(finite, p, torch.zeros_like(p)).double().sum(), finite.double().sum(),
                            s.double().mean()]).tolist()                                   # the one sync
        return {"psnr": both[0] / both[1] if both[1] > 0 else 0.0, "ssim": both[2]}

    def __str__(self) -> str:
        m = self.get_metrics()
        return f"PSNR: {m['psnr']:.2f} dB, SSIM: {m['ssim']:.4f}"
