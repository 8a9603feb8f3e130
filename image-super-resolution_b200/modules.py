"""Parameter holders for the fusion network.

The CUDA pipeline (``pipeline.py``) reads raw ``.weight/.bias/.running_*``
tensors; these classes exist so that ``state_dict()`` has exactly the reference's
226 keys, default initialisation under a given
``torch.manual_seed`` draws the RNG in the reference's order, and EMA / AdamW /
``clip_grad_norm_`` see ordinary leaf ``nn.Parameter``s.  None of them computes
anything: every ``forward`` raises, the work happens in the sm_100a kernels.

Layout being mirrored (construction order matters for RNG parity):
  freq_decomp      src/models/multi_domain_frequency.py:66-385, 533-576
  cross_band       src/models/large_kernel_attention.py:38-149, 156-205
  collaborative    src/models/large_kernel_attention.py:251-322
  multi_res        src/models/hierarchical_fusion.py:25-129
  dynamic_selector src/models/enhanced_fusion_v2.py:417-448
  edge_enhance     src/models/edge_enhancement.py:36-180
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

# db4 analysis low-pass taps (multi_domain_frequency.py:39-48); high-pass is the QMF
# mirror: hi[i] = (-1)^(i+1) * lo[7-i]  (:50-59).
_DB4_LO = (-0.010597401784997278, 0.032883011666982945, 0.030841381835986965,
           -0.18703481171888114, -0.027983769416983849, 0.63088076792959036,
           0.71484657055291582, 0.23037781330885523)
_DB4_HI = tuple(((-1.0) ** (i + 1)) * _DB4_LO[7 - i] for i in range(8))


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - never a compute path
        raise RuntimeError(f"{type(self).__name__} is a parameter holder; "
                           "compute runs in the sm_100a kernels of CompleteEnhancedFusionSR")


def zigzag_order(n: int) -> torch.Tensor:
    """JPEG zigzag index of every (row, col) of an n x n block."""
    idx = torch.zeros(n, n, dtype=torch.long)
    k = 0
    for s in range(2 * n - 1):
        rows = range(min(s, n - 1), max(0, s - n + 1) - 1, -1) if s % 2 == 0 \
            else range(max(0, s - n + 1), min(s, n - 1) + 1)
        for i in rows:
            idx[i, s - i] = k
            k += 1
    return idx


class DCTBands(_Holder):
    def __init__(self, block_size: int = 8):
        super().__init__()
        n = block_size
        self.block_size = n
        basis = torch.zeros(n, n)
        for k in range(n):
            for j in range(n):
                # float64 numpy-style evaluation then rounding to fp32, as the reference does
                basis[k, j] = math.sqrt(1.0 / n) if k == 0 else \
                    math.sqrt(2.0 / n) * math.cos(math.pi * k * (2 * j + 1) / (2 * n))
        self.register_buffer("dct_basis", basis)
        self.register_buffer("dct_basis_t", basis.T)
        zz = zigzag_order(n)
        lo_t, hi_t = (n * n) // 3, 2 * (n * n) // 3
        self.register_buffer("low_mask", (zz < lo_t).float())
        self.register_buffer("mid_mask", ((zz >= lo_t) & (zz < hi_t)).float())
        self.register_buffer("high_mask", (zz >= hi_t).float())
        self.band_scale = nn.Parameter(torch.ones(3))


class DWTBands(_Holder):
    def __init__(self, in_channels: int = 3):
        super().__init__()
        self.in_channels = in_channels
        lo = torch.tensor(_DB4_LO, dtype=torch.float32)
        hi = torch.tensor(_DB4_HI, dtype=torch.float32)
        self.register_buffer("lo_row", lo.reshape(1, 1, 1, 8).repeat(in_channels, 1, 1, 1))
        self.register_buffer("hi_row", hi.reshape(1, 1, 1, 8).repeat(in_channels, 1, 1, 1))
        self.register_buffer("lo_col", lo.reshape(1, 1, 8, 1).repeat(in_channels, 1, 1, 1))
        self.register_buffer("hi_col", hi.reshape(1, 1, 8, 1).repeat(in_channels, 1, 1, 1))
        self.subband_scale = nn.Parameter(torch.ones(4))


class FFTBands(_Holder):
    def __init__(self, init_mask_size: int = 64):
        super().__init__()
        ax = torch.linspace(-1, 1, init_mask_size)
        yy, xx = torch.meshgrid(ax, ax, indexing="ij")
        logits = 3.0 * (0.5 - torch.sqrt(xx ** 2 + yy ** 2))
        self.freq_mask_logits = nn.Parameter(logits[None, None])
        self.temperature = nn.Parameter(torch.tensor(5.0))
        self.band_scale = nn.Parameter(torch.ones(2))


class FrequencyBands(_Holder):
    """freq_decomp.{dct,dwt,fft}; the 9->3 band-fusion module is disabled upstream
    (enhanced_fusion_v2.py:528-531) so it has no parameters here."""

    def __init__(self, block_size: int = 8, in_channels: int = 3, fft_mask_size: int = 64):
        super().__init__()
        self.dct = DCTBands(block_size)
        self.dwt = DWTBands(in_channels)
        self.fft = FFTBands(fft_mask_size)


class LKA(_Holder):
    def __init__(self, dim: int, k: int = 21):
        super().__init__()
        self.local_conv = nn.Conv2d(dim, dim, 5, padding=2, groups=dim, bias=False)
        self.h_conv = nn.Conv2d(dim, dim, (1, k), padding=(0, k // 2), groups=dim, bias=False)
        self.v_conv = nn.Conv2d(dim, dim, (k, 1), padding=(k // 2, 0), groups=dim, bias=False)
        self.pw_conv = nn.Conv2d(dim, dim, 1, bias=False)
        self.bn = nn.BatchNorm2d(dim)
        for c in (self.local_conv, self.h_conv, self.v_conv):
            nn.init.kaiming_normal_(c.weight, mode="fan_out")
        nn.init.xavier_uniform_(self.pw_conv.weight)


class LKABlockParams(_Holder):
    def __init__(self, dim: int, k: int = 21, ffn_ratio: float = 2.0):
        super().__init__()
        self.norm1 = nn.BatchNorm2d(dim)
        self.lka = LKA(dim, k)
        self.norm2 = nn.BatchNorm2d(dim)
        hid = int(dim * ffn_ratio)
        self.ffn = nn.Sequential(nn.Conv2d(dim, hid, 1), nn.GELU(), nn.Conv2d(hid, dim, 1))
        self.scale1 = nn.Parameter(torch.tensor(0.1))
        self.scale2 = nn.Parameter(torch.tensor(0.1))


class CrossBand(_Holder):
    def __init__(self, dim: int = 64, num_bands: int = 9, num_heads: int = 4, lka_kernel: int = 21):
        super().__init__()
        self.num_bands, self.dim, self.num_heads = num_bands, dim, num_heads
        self.band_proj = nn.Conv2d(3, dim, 1)
        self.band_attention = nn.MultiheadAttention(dim, num_heads, batch_first=True, dropout=0.1)
        self.norm = nn.LayerNorm(dim)
        self.lka_block = LKABlockParams(dim, lka_kernel, 2.0)
        self.out_proj = nn.Conv2d(dim, 3, 1)


class Collaborative(_Holder):
    def __init__(self, num_experts: int = 4, feature_dim: int = 128, num_heads: int = 8,
                 lka_kernel: int = 21):
        super().__init__()
        self.num_experts, self.feature_dim, self.num_heads = num_experts, feature_dim, num_heads
        self.align_layers = nn.ModuleDict({
            "drct": nn.Conv2d(180, feature_dim, 1),
            "grl": nn.Conv2d(180, feature_dim, 1),
            "nafnet": nn.Conv2d(64, feature_dim, 1),
            "mamba": nn.Conv2d(180, feature_dim, 1),
        })
        self.cross_attn = nn.MultiheadAttention(feature_dim, num_heads, batch_first=True, dropout=0.1)
        self.norm1 = nn.LayerNorm(feature_dim)
        self.norm2 = nn.LayerNorm(feature_dim)
        self.ffn = nn.Sequential(nn.Linear(feature_dim, feature_dim * 2), nn.GELU(),
                                 nn.Linear(feature_dim * 2, feature_dim))
        self.lka_global = LKABlockParams(feature_dim, lka_kernel, 2.0)
        self.modulation = nn.ModuleList([
            nn.Sequential(nn.Conv2d(feature_dim, feature_dim // 4, 1), nn.GELU(),
                          nn.Conv2d(feature_dim // 4, 3, 1), nn.Sigmoid())
            for _ in range(num_experts)])


class _Gate(_Holder):
    def __init__(self, ch: int):
        super().__init__()
        self.gate = nn.Sequential(nn.Conv2d(ch, ch // 4, 1), nn.GELU(), nn.Conv2d(ch // 4, 1, 1), nn.Sigmoid())


class _Res(_Holder):
    def __init__(self, ch: int):
        super().__init__()
        self.block = nn.Sequential(nn.Conv2d(ch, ch, 3, 1, 1, bias=False), nn.GELU(),
                                   nn.Conv2d(ch, ch, 3, 1, 1, bias=False))
        self.scale = nn.Parameter(torch.tensor(0.1))


class HierarchicalFusion(_Holder):
    def __init__(self, num_experts: int = 4, base_channels: int = 128):
        super().__init__()
        self.num_experts, self.base_channels = num_experts, base_channels
        cin, c = num_experts * 3, base_channels

        def pair(i, m, o):
            return nn.Sequential(nn.Conv2d(i, m, 3, 1, 1), nn.GELU(), nn.Conv2d(m, o, 3, 1, 1), nn.GELU())

        self.stage1_conv = pair(cin, c, c)
        self.stage1_gate = _Gate(c)
        self.stage1_res = _Res(c)
        self.stage2_conv = pair(c + cin, c, c)
        self.stage2_gate = _Gate(c)
        self.stage2_res = _Res(c)
        self.stage3_conv = pair(c + cin, c, c // 2)
        self.stage3_gate = _Gate(c // 2)
        self.stage3_res = _Res(c // 2)
        self.to_rgb = nn.Sequential(nn.Conv2d(c // 2, c // 4, 3, 1, 1), nn.GELU(), nn.Conv2d(c // 4, 3, 3, 1, 1))
        self.residual_weight_1_2 = nn.Parameter(torch.tensor(0.2))
        self.residual_weight_2_3 = nn.Parameter(torch.tensor(0.2))


class DynamicExpertSelector(_Holder):
    """Per-pixel difficulty / expert gate nets (enhanced_fusion_v2.py:417-448)."""

    def __init__(self, in_channels: int = 3, hidden_dim: int = 32, num_experts: int = 4):
        super().__init__()
        self.num_experts = num_experts
        h = hidden_dim
        self.difficulty_net = nn.Sequential(
            nn.Conv2d(in_channels, h, 3, 1, 1), nn.ReLU(inplace=True),
            nn.Conv2d(h, h, 3, 1, 1), nn.ReLU(inplace=True),
            nn.Conv2d(h, 1, 3, 1, 1), nn.Sigmoid())
        self.gate_net = nn.Sequential(
            nn.Conv2d(in_channels, h, 3, 1, 1), nn.ReLU(inplace=True),
            nn.Conv2d(h, h, 3, 1, 1), nn.ReLU(inplace=True),
            nn.Conv2d(h, num_experts, 1))
        self.temperature = nn.Parameter(torch.tensor(10.0))


class _Blur(_Holder):
    def __init__(self, channels: int = 3, k: int = 5, sigma: float = 1.5):
        super().__init__()
        ax = torch.arange(k, dtype=torch.float32) - k // 2
        g = torch.exp(-(ax ** 2) / (2 * sigma ** 2))
        g = g / g.sum()
        self.register_buffer("kernel", (g[:, None] * g[None, :]).expand(channels, 1, k, k).contiguous())


class _EdgeAttn(_Holder):
    def __init__(self, ch: int):
        super().__init__()
        self.attn = nn.Sequential(nn.Conv2d(ch, ch // 4, 1), nn.GELU(), nn.Conv2d(ch // 4, 1, 3, 1, 1), nn.Sigmoid())


class _EdgeRefiner(_Holder):
    def __init__(self, in_ch: int = 3, feat: int = 32):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, feat, 3, 1, 1)
        self.conv2 = nn.Conv2d(feat, feat, 3, 1, 1)
        self.conv3 = nn.Conv2d(feat, feat, 3, 1, 1)
        self.act = nn.GELU()
        self.proj = nn.Conv2d(in_ch, feat, 1) if in_ch != feat else nn.Identity()
        self.attn = _EdgeAttn(feat)


class LaplacianEdge(_Holder):
    def __init__(self, num_levels: int = 3, channels: int = 32, edge_strength: float = 0.15):
        super().__init__()
        self.num_levels, self.channels = num_levels, channels
        self.gaussian = _Blur(3, 5, 1.5)
        self.edge_refiners = nn.ModuleList([_EdgeRefiner(3, channels) for _ in range(num_levels)])
        self.fusion = nn.Sequential(nn.Conv2d(num_levels * channels, channels, 3, 1, 1), nn.GELU(),
                                    nn.Conv2d(channels, 3, 3, 1, 1))
        self.level_weights = nn.Parameter(torch.ones(num_levels) / num_levels)
        self.edge_gate = nn.Sequential(nn.Conv2d(6, 16, 3, 1, 1), nn.GELU(), nn.Conv2d(16, 1, 3, 1, 1), nn.Sigmoid())
        self.edge_strength = nn.Parameter(torch.tensor(edge_strength))
