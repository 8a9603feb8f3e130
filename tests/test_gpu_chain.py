"""Tile-resident (fused chain) tcgen05 kernels against the layer-by-layer tcgen05 path they replace, on the same bf16
inputs and weights: the two differ only by fp32 summation order and by where bf16 rounding of side inputs happens."""
import pytest
import torch

import isr_b200
from isr_b200 import _cabi as K
from oracle import fusion_oracle as O
from oracle.perturb import perturb_state_dict

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _model(dev):
    torch.manual_seed(0)
    m = isr_b200.CompleteEnhancedFusionSR(None)
    m.load_state_dict(perturb_state_dict(m.state_dict(), seed=3))
    m.eval().to(dev)
    m.precision = "bf16"
    return m


@pytest.mark.parametrize("B,H,W", [(1, 16, 16), (2, 13, 29), (1, 40, 56)])
def test_edge_refiner_chain_matches_layer_by_layer(B, H, W):
    """ffsr_edge_refiner_chain (csrc/edge_chain.cu) vs proj / conv1 / conv2 / conv3 / attn0 / attn2 as k_conv_tc launches
    + ffsr_edge_attn_upsample: the 96-channel concat of the three refined pyramid levels and the final image."""
    dev = _cuda()
    m = _model(dev)
    lr, imgs, fts, _ = O.synthetic_inputs(B, H, W)
    lr, imgs, fts = lr.to(dev), {k: v.to(dev) for k, v in imgs.items()}, {k: v.to(dev) for k, v in fts.items()}
    with torch.no_grad():
        m.forward_with_precomputed(lr, imgs, fts)
        eng = m._engine
        assert eng.edge_chain, "FFSR_EDGE_CHAIN0 is set: nothing to compare"
        sr_chain = m.forward_with_precomputed(lr, imgs, fts).float().cpu()
        cat_chain = eng.workspace("ee.cat96").float().cpu().clone()
        eng.edge_chain = False
        sr_ref = m.forward_with_precomputed(lr, imgs, fts).float().cpu()
        cat_ref = eng.workspace("ee.cat96").float().cpu().clone()
        eng.edge_chain = True
    scale = cat_ref.abs().max().item()
    for lv in range(3):
        a, b = cat_chain[..., 32 * lv:32 * lv + 32], cat_ref[..., 32 * lv:32 * lv + 32]
        err = (a - b).abs().max().item()
        # bf16 storage of three chained 32-channel layers: a few bf16 ulps of the feature scale
        assert err <= 0.04 * max(scale, 1e-3), f"level {lv}: max-abs {err:.4e} (feature scale {scale:.3e})"
        assert torch.nn.functional.cosine_similarity(a.reshape(-1), b.reshape(-1), dim=0).item() > 0.9995, lv
    assert (sr_chain - sr_ref).abs().max().item() <= 5e-3
