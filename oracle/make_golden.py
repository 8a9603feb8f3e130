"""Generate golden fixtures by running the REFERENCE ITSELF (build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Imports ``/root/reference`` (read-only, never copied), builds
``CompleteEnhancedFusionSR(None)`` under ``torch.manual_seed(0)``, feeds the
SURVEY §8(d) synthetic inputs and stores the outputs + per-phase intermediates
under ``tests/golden/``.  The reference tree does not travel to the GPU box, the
fixtures do.  Test infrastructure only.

Fixtures:
  state_hashes.json        sha256 of every state_dict tensor at seed 0 (pins that OUR
                           module initialises bit-identically to the reference)
  case_default_16x16.npz   default-init weights, B=1, 16x16 LR, eval
  case_perturbed_17x23.npz perturbed weights (oracle.perturb_state_dict), B=1, 17x23 LR
                           (odd sizes: DCT reflect pad, DWT odd lengths, rfft odd W), eval
  case_perturbed_b2_9x11.npz  perturbed weights, B=2, 9x11, eval, SR + gates only
  case_nofeat_16x24.npz    expert_feats=None (Phase 4 skipped), eval
  case_train_b2_12x12.npz  train mode, attention dropout forced to 0, BN running-stat updates
  losses_24x24.npz         stage-3 loss components on a fixed (sr, hr) pair
"""
import hashlib
import json
import math
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from oracle.fusion_oracle import EXPERT_ORDER, synthetic_inputs  # noqa: E402
from oracle.perturb import perturb_state_dict  # noqa: E402

# pywt is absent offline; SWTLoss only reads the four Haar taps from it
_s = 1 / math.sqrt(2)
_fake = types.ModuleType("pywt")
_fake.Wavelet = lambda name: types.SimpleNamespace(dec_lo=[_s, _s], dec_hi=[-_s, _s])
sys.modules.setdefault("pywt", _fake)

from src.models.enhanced_fusion_v2 import CompleteEnhancedFusionSR  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def build(perturbed: bool):
    torch.manual_seed(0)
    m = CompleteEnhancedFusionSR(expert_ensemble=None)
    if perturbed:
        m.load_state_dict(perturb_state_dict(m.state_dict(), seed=7), strict=True)
    return m


def run_case(m, B, H, W, feats=True, training=False):
    lr, imgs, fts, hr = synthetic_inputs(B, H, W, seed=1234, feats=feats)
    if training:
        m.train()
        m.cross_band.band_attention.dropout = 0.0
        m.collaborative.cross_attn.dropout = 0.0
    else:
        m.eval()
    with torch.no_grad():
        sr, ints = m._run_pipeline(lr, [imgs[k] for k in EXPERT_ORDER], fts or {}, 4 * H, 4 * W, {}, True)
        sr2 = m.forward_with_precomputed(lr, imgs, fts) if not training else sr
    assert training or torch.equal(sr, sr2)
    return sr, ints


def pack(sr, ints, full=True):
    d = {"sr": sr.numpy()}
    d["gates"] = ints["gates"].numpy()
    d["difficulty"] = ints["difficulty"].numpy()
    if full:
        d["raw_9_bands"] = torch.stack(ints["raw_9_bands"], 1).numpy()
        d["enhanced_9_bands"] = torch.stack(ints["enhanced_9_bands"], 1).numpy()
        d["routing_lr"] = ints["routing_lr"].numpy()
        d["fused_before_dynamic"] = ints["fused_before_dynamic"].numpy()
        if "collaborative_outputs" in ints:
            d["collaborative_outputs"] = torch.stack(ints["collaborative_outputs"], 1).numpy()
    return d


def main():
    torch.set_num_threads(8)
    m0 = build(False)
    with open(os.path.join(OUT, "state_hashes.json"), "w") as f:
        json.dump({k: [list(v.shape), str(v.dtype), sha(v)] for k, v in m0.state_dict().items()}, f, indent=0)

    sr, ints = run_case(m0, 1, 16, 16)
    np.savez_compressed(os.path.join(OUT, "case_default_16x16.npz"), **pack(sr, ints))

    mp = build(True)
    sr, ints = run_case(mp, 1, 17, 23)
    np.savez_compressed(os.path.join(OUT, "case_perturbed_17x23.npz"), **pack(sr, ints))

    sr, ints = run_case(mp, 2, 9, 11)
    np.savez_compressed(os.path.join(OUT, "case_perturbed_b2_9x11.npz"), **pack(sr, ints, full=False))

    sr, ints = run_case(mp, 1, 16, 24, feats=False)
    np.savez_compressed(os.path.join(OUT, "case_nofeat_16x24.npz"), **pack(sr, ints, full=False))

    mt = build(True)
    before = {k: v.clone() for k, v in mt.state_dict().items()}
    sr, ints = run_case(mt, 2, 12, 12, training=True)
    d = pack(sr, ints, full=False)
    for k, v in mt.state_dict().items():
        if ("running_" in k or "num_batches" in k) and not torch.equal(v, before[k]):
            d["bn::" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "case_train_b2_12x12.npz"), **d)

    # stage-3 losses (SURVEY Appendix D recipe 6) on a fixed pair
    from src.losses.perceptual_loss import L1Loss, SSIMLoss, FFTLoss, SWTLoss
    g = torch.Generator().manual_seed(99)
    a = torch.rand(2, 3, 24, 24, generator=g)
    b = (a + 0.1 * torch.randn(2, 3, 24, 24, generator=g)).clamp(0, 1)
    comps = {
        "l1": L1Loss()(a, b), "ssim": SSIMLoss()(a, b), "fft": FFTLoss(focus_high_freq=True)(a, b),
        "swt": SWTLoss(wavelet="haar", level=2, use_gpu_approximation=True)(a, b),
    }
    np.savez_compressed(os.path.join(OUT, "losses_24x24.npz"), pred=a.numpy(), target=b.numpy(),
                        **{k: np.float64(v.item()) for k, v in comps.items()})
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
