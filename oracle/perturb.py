"""Deterministic weight perturbation for parity tests (TEST INFRASTRUCTURE ONLY).

Default initialisation leaves many tensors at trivial values (BN running stats 0/1,
attention biases 0, band scales 1), which would hide a missing bias / scale / BN fold
in a kernel.  ``perturb_state_dict`` moves every learnable tensor and every BatchNorm
running statistic off its default while leaving the fixed transform constants (DCT
basis and masks, db4 taps, Gaussian kernel) untouched.
"""
import torch

_FIXED = ("dct_basis", "dct_basis_t", "low_mask", "mid_mask", "high_mask",
          "lo_row", "hi_row", "lo_col", "hi_col", "gaussian.kernel")


def perturb_state_dict(sd, seed: int = 7):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if not torch.is_floating_point(v) or k.endswith(_FIXED):
            out[k] = v.clone()
            continue
        n = torch.randn(v.shape, generator=g, dtype=torch.float32)
        if k.endswith("running_var"):
            out[k] = (v + 0.5 * n.abs()).to(v.dtype)
        elif k.endswith("running_mean"):
            out[k] = (v + 0.1 * n).to(v.dtype)
        elif v.dim() == 0:
            out[k] = (v * (1.0 + 0.1 * n)).to(v.dtype)          # scalar scales / temperatures
        else:
            scale = 0.05 if v.dim() > 1 else 0.1
            out[k] = (v + scale * max(float(v.abs().mean()), 0.05) * n).to(v.dtype)
    return out
