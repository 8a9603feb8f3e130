"""CPU-only checks of the host side: C-ABI library loads and exports every declared symbol,
the drop-in module mirrors the reference interface, and the kernel orchestration is
well-formed (dry run against a recording fake of the library: no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest
import torch

import isr_b200
from isr_b200 import _cabi
from isr_b200.pipeline import FusionEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "ffsr_b200.h")).read()
    declared = set(re.findall(r"\b(ffsr_[a-z0-9_]+)\s*\(", header))
    declared.discard("ffsr_conv_params")
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ffsr_b200.h but not exported"
    assert declared == set(_cabi.PROTOTYPES), declared ^ set(_cabi.PROTOTYPES)
    assert b"sm_100a" in lib.ffsr_version()


def test_conv_params_struct_matches_header_layout():
    # 64-bit: 5 ptr/ll + 6 int + ... ; the C side is compiled from the same field order, so a
    # size check catches accidental drift between _cabi.ConvParams and ffsr_conv_params.
    assert C.sizeof(_cabi.ConvParams) == _cabi.load().ffsr_conv_params_size() == 280


def test_module_interface_matches_reference():
    torch.manual_seed(0)
    m = isr_b200.CompleteEnhancedFusionSR(None)
    assert m.cached_mode is True and m.num_experts == 4 and m.upscale == 4
    assert m.get_trainable_params() == 1_433_217 and m.get_frozen_params() == 0
    assert all(m.get_improvement_status().values())
    with pytest.raises(RuntimeError, match="cached mode"):
        m(torch.rand(1, 3, 16, 16))
    # no silent CPU fallback: CPU tensors are refused loudly
    m.eval()
    imgs = {k: torch.rand(1, 3, 64, 64) for k in isr_b200.EXPERT_ORDER}
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.forward_with_precomputed(torch.rand(1, 3, 16, 16), imgs, None)
    # strict state_dict round trip with itself and the factory
    m2 = isr_b200.create_enhanced_fusion(None, {"refine_depth": 6})
    m2.load_state_dict(m.state_dict(), strict=True)
    assert "trainable=1,433,217" in repr(m2)


class _FakeLib:
    """Records call names; implements only the two size queries."""

    def __init__(self):
        self.calls = []

    def ffsr_dwt_sub_size(self, H, W, hs, ws):
        hs._obj.value = (H + 6) // 2 + 1
        ws._obj.value = (W + 6) // 2 + 1
        return 0

    def ffsr_fft_workspace_bytes(self, B, H, W):
        wf = W // 2 + 1
        return 2 * B * 3 * H * wf * 16 + (H * wf * 4 + 255) // 256 * 256

    def __getattr__(self, name):
        def f(*args):
            self.calls.append(name)
            return 0
        f.__name__ = name
        return f


@pytest.mark.parametrize("with_feats,want_inter,precision", [(True, False, "fp32"), (True, True, "fp32"),
                                                             (False, False, "fp32"), (True, False, "bf16")])
def test_orchestration_dry_run(with_feats, want_inter, precision):
    torch.manual_seed(0)
    m = isr_b200.CompleteEnhancedFusionSR(None).eval()
    m.precision = precision
    eng = FusionEngine(m)
    eng.lib = _FakeLib()
    eng._get_stream = lambda dev: C.c_void_p(0)
    eng._sm_count = lambda dev: 148
    B, H, W = 2, 12, 20
    lr = torch.rand(B, 3, H, W)
    imgs = [torch.rand(B, 3, 4 * H, 4 * W) for _ in range(4)]
    feats = {k: torch.randn(B, 64 if k == "nafnet" else 180, H, W) for k in isr_b200.EXPERT_ORDER} if with_feats else {}
    out, inter = eng._forward(lr, imgs, feats, B, H, W, want_inter)
    assert out.shape == (B, 3, 4 * H, 4 * W)
    calls = eng.lib.calls
    # conv launches: the Phase-3 LKA tail (3 convs) is one ffsr_lka_tail64 launch in both modes; bf16 mode also has one
    # grouped align conv instead of four, each edge refiner (6 convs x 3 levels) as one ffsr_edge_refiner_chain launch and
    # the Phase-4 LKA tail + modulation layer 0 (4 convs) as one ffsr_lka_tail128_mod launch, and the Phase-4 token pipeline
    # (qkv, out, ffn0, ffn2 convs + two LayerNorms + the attention core) as ffsr_token_attn_chain + ffsr_token_ffn_chain
    # Phase 6 (six selector convs + the gate normalisation) is one ffsr_selector_fused launch in both modes
    expect = (63 if with_feats else 51) - 3 - 6
    if precision == "bf16":
        expect -= 3 + 18 + 4 + 4 + (1 if with_feats else 0)     # + the grouped align conv, now inside ffsr_align_tokens
    assert calls.count("ffsr_conv2d") == expect, calls.count("ffsr_conv2d")
    assert calls.count("ffsr_lka_tail64") == 1
    assert calls.count("ffsr_selector_fused") == 1 and "ffsr_gate_finalize" not in calls
    assert calls.count("ffsr_edge_refiner_chain") == (3 if precision == "bf16" else 0)
    assert calls.count("ffsr_lka_tail128_mod") == (1 if (precision == "bf16" and with_feats) else 0)
    chain = precision == "bf16" and with_feats
    assert ("ffsr_token_attention" in calls) == (with_feats and not chain)
    assert calls.count("ffsr_token_attn_chain") == calls.count("ffsr_token_ffn_chain") == (1 if chain else 0)
    assert calls[-1] == "ffsr_final_combine"
    if want_inter:
        assert set(inter) >= {"raw_9_bands", "enhanced_9_bands", "routing_lr", "collaborative_outputs",
                              "fused_before_dynamic", "gates", "difficulty"}
        assert len(inter["raw_9_bands"]) == 9 and inter["gates"].shape == (B, 4, H, W)


def test_tile_grid_covers_the_image_on_the_8px_grid():
    from isr_b200.serving import tile_grid, TILE_HALO_LR
    for (H, W, ty, tx) in [(339, 510, 1, 2), (339, 510, 2, 4), (64, 64, 2, 2), (17, 200, 1, 8)]:
        t = tile_grid(H, W, ty, tx)
        assert len(t) == ty * tx
        cover = set()
        for y0, y1, x0, x1 in t:
            assert 0 <= y0 < y1 <= H and 0 <= x0 < x1 <= W and y0 % 8 == 0 and x0 % 8 == 0
            cover.update((y, x) for y in range(y0, y1, 7) for x in range(x0, x1, 7))
        assert sum((y1 - y0) * (x1 - x0) for y0, y1, x0, x1 in t) == H * W
    assert TILE_HALO_LR % 8 == 0 and TILE_HALO_LR * 4 >= 110        # receptive field behind the bands: <= 110 HR px
    import pytest
    with pytest.raises(ValueError):
        tile_grid(16, 16, 4, 1)


def test_device_only_entry_points_refuse_cpu_tensors():
    import pytest
    import torch
    from isr_b200 import metrics as MT
    from isr_b200.serving import fuse_tiled
    a = torch.rand(1, 3, 16, 16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        MT.psnr_ssim_per_image(a, a)
    m = isr_b200.CompleteEnhancedFusionSR(None).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        fuse_tiled(m, torch.rand(1, 3, 16, 16), {k: torch.rand(1, 3, 64, 64) for k in ("drct", "grl", "nafnet", "mamba")}, None)


def _unplane(planes: torch.Tensor, n: int, k: int) -> torch.Tensor:
    """Inverse of pipeline._planes_bf16: [k/8][n][8] -> [n][k] (fp32)."""
    return planes.float().view(k // 8, n, 8).permute(1, 0, 2).reshape(n, k)


def test_token_chain_packing_is_layernorm_folded_into_the_linear():
    """pack_token_attn / pack_token_ffn (csrc/token_chain.cu operands): with g = gamma o W rounded to bf16,
    rstd * (x g^T - mean * colsum(g)) + (W beta + b) must equal Linear(LayerNorm(x)) up to the bf16 rounding of g, the q rows
    carry the 1/sqrt(head_dim) = 1/4 of the scores, and the planes decode back to the matrices."""
    from isr_b200.pipeline import pack_token_attn, pack_token_ffn
    torch.manual_seed(1)
    m = isr_b200.CompleteEnhancedFusionSR(None).eval()
    co = m.collaborative
    with torch.no_grad():
        co.norm1.weight.uniform_(0.5, 1.5); co.norm1.bias.normal_(0, 0.2)
        co.norm2.weight.uniform_(0.5, 1.5); co.norm2.bias.normal_(0, 0.2)
    wa, pa = pack_token_attn(co)
    g = _unplane(wa[:384 * 128], 384, 128).double()
    wo = _unplane(wa[384 * 128:], 128, 128)
    cs, bq, bo = pa[:384].double(), pa[384:768].double(), pa[768:]
    assert torch.equal(wo, co.cross_attn.out_proj.weight.detach().to(torch.bfloat16).float())
    assert torch.equal(bo, co.cross_attn.out_proj.bias.detach())
    assert torch.allclose(cs, g.sum(1), atol=1e-5)
    x = torch.randn(64, 128, dtype=torch.float64) * 0.8 + 0.3
    mean, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
    rstd = (var + 1e-5).rsqrt()
    got = rstd * (x @ g.t() - mean * cs[None, :]) + bq[None, :]
    ln = torch.nn.functional.layer_norm(x, (128,), co.norm1.weight.double(), co.norm1.bias.double(), 1e-5)
    want = ln @ co.cross_attn.in_proj_weight.double().t() + co.cross_attn.in_proj_bias.double()
    want[:, :128] *= 0.25
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()          # bf16 rounding of g only
    wf, pf = pack_token_ffn(co)
    g0 = _unplane(wf[:256 * 128], 256, 128).double()
    w2 = _unplane(wf[256 * 128:], 128, 256)
    assert torch.equal(w2, co.ffn[2].weight.detach().to(torch.bfloat16).float())
    got = rstd * (x @ g0.t() - mean * pf[:256].double()[None, :]) + pf[256:512].double()[None, :]
    ln2 = torch.nn.functional.layer_norm(x, (128,), co.norm2.weight.double(), co.norm2.bias.double(), 1e-5)
    want = ln2 @ co.ffn[0].weight.double().t() + co.ffn[0].bias.double()
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()
    assert torch.equal(pf[512:], co.ffn[2].bias.detach())


def test_selector_and_align_blobs_follow_the_kernel_layouts():
    from isr_b200.pipeline import pack_selector, pack_align_tokens, EXPERT_ORDER
    torch.manual_seed(2)
    m = isr_b200.CompleteEnhancedFusionSR(None).eval()
    ds = m.dynamic_selector
    blob = pack_selector(ds)
    # csrc/selector.cu: d0w(864) d0b(32) d2w(9216) d2b(32) d4w(288) d4b(1+3) g0w g0b g2w g2b g4w(128) g4b(4)
    assert blob.numel() == 2 * (864 + 32 + 9216 + 32) + 288 + 4 + 128 + 4
    o = 864 + 32
    w2 = blob[o:o + 9216].view(9, 32, 32)                               # [tap][ci][co]
    assert torch.equal(w2[4], ds.difficulty_net[2].weight.detach()[:, :, 1, 1].t().contiguous())
    assert torch.equal(blob[o + 9216:o + 9216 + 32], ds.difficulty_net[2].bias.detach())
    assert torch.equal(blob[-4:], ds.gate_net[4].bias.detach())
    wb, bb = pack_align_tokens(m.collaborative)
    assert wb.numel() == 4 * 24 * 128 * 8 and tuple(bb.shape) == (4, 128)
    for e, n in enumerate(EXPERT_ORDER):
        l = m.collaborative.align_layers[n]
        w = _unplane(wb[e * 24 * 128 * 8:(e + 1) * 24 * 128 * 8], 128, 192)
        cin = l.weight.shape[1]
        assert torch.equal(w[:, :cin], l.weight.detach().reshape(128, cin).to(torch.bfloat16).float())
        assert torch.count_nonzero(w[:, cin:]) == 0                      # K zero-padded to 192
