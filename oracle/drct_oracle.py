"""CPU oracle for the DRCT-L expert forward (TEST INFRASTRUCTURE ONLY; SURVEY §8f N1 -- prepared ahead of its CUDA path).

Functional restatement, over a plain ``state_dict``, of ``src/models/drct/drct_arch.py`` as configured by
``create_drct_model`` (``src/models/drct/__init__.py:84-131``: DRCT-L = embed 180, 12 RDGs, window 16, mlp_ratio 2,
pixelshuffle x4, ``resi_connection='1conv'``, growth 32): shallow conv, patch-embed LayerNorm, RDG blocks of five
(shifted-)window attention blocks with dense growth and 1x1 "adjust" convs, final norm, ``conv_after_body`` (whose
output is the cached 180-channel feature, ``src/models/expert_loader.py`` hook) + skip, pixel-shuffle reconstruction.
Eval semantics only (dropout / drop-path are identities).

Parity pin: ``tests/golden/drct_small.npz`` holds the output and the hook feature of the REFERENCE class itself on a
reduced configuration (2 RDGs, window 8, every channel count of DRCT-L) with weights synthesised from a seed by
``synth_state_dict`` (``oracle/make_drct_golden.py``); ``tests/test_drct_oracle.py`` checks this restatement against
it, and against the reference class directly when ``/root/reference`` is present.  All citations: ``/root/reference``.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

RGB_MEAN = (0.4488, 0.4371, 0.4040)          # drct_arch.py:665-666


# ---------------------------------------------------------------------------------------------------
# shapes and synthetic weights (so that goldens need no checkpoint)
# ---------------------------------------------------------------------------------------------------
def relative_position_index(ws: int) -> torch.Tensor:
    """WindowAttention.__init__, drct_arch.py:153-165."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    flat = torch.flatten(coords, 1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def swin_dims(embed_dim: int, num_heads: int, gc: int, mlp_ratio: float):
    """(dim, heads, mlp_hidden, shifted) of swin1..5 of one RDG, drct_arch.py:230-277."""
    out = []
    for j in range(5):
        d = embed_dim + j * gc
        h = num_heads - (d % num_heads) if j else num_heads
        r = mlp_ratio if j < 3 else 1
        out.append((d, h, int(d * r), j % 2 == 1))
    return out


def state_shapes(embed_dim=180, n_rdg=12, window=16, num_heads=6, gc=32, mlp_ratio=2, num_feat=64, img_size=64
                 ) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """name -> (shape, kind) of ``DRCT(...).state_dict()`` in its order (kind: 'float' | 'index' | 'mask')."""
    s: Dict[str, Tuple[Tuple[int, ...], str]] = {}

    def lin(p, o, i):
        s[p + ".weight"], s[p + ".bias"] = ((o, i), "float"), ((o,), "float")

    def conv(p, o, i, k):
        s[p + ".weight"], s[p + ".bias"] = ((o, i, k, k), "float"), ((o,), "float")

    conv("conv_first", embed_dim, 3, 3)
    lin_ln = lambda p, d: s.update({p + ".weight": ((d,), "float"), p + ".bias": ((d,), "float")})      # noqa: E731
    lin_ln("patch_embed.norm", embed_dim)
    nw = (img_size // window) ** 2
    for i in range(n_rdg):
        for j, (d, h, hid, shifted) in enumerate(swin_dims(embed_dim, num_heads, gc, mlp_ratio)):
            p = f"layers.{i}.swin{j + 1}"
            if shifted:
                s[p + ".attn_mask"] = ((nw, window * window, window * window), "mask")
            lin_ln(p + ".norm1", d)
            s[p + ".attn.relative_position_bias_table"] = (((2 * window - 1) ** 2, h), "float")
            s[p + ".attn.relative_position_index"] = ((window * window, window * window), "index")
            lin(p + ".attn.qkv", 3 * d, d)
            lin(p + ".attn.proj", d, d)
            lin_ln(p + ".norm2", d)
            lin(p + ".mlp.fc1", hid, d)
            lin(p + ".mlp.fc2", d, hid)
            conv(f"layers.{i}.adjust{j + 1}", gc if j < 4 else embed_dim, d, 1)
        # module registration order inside RDG is swin1, adjust1, swin2, adjust2, ... (drct_arch.py:230-277): handled above
    lin_ln("norm", embed_dim)
    conv("conv_after_body", embed_dim, embed_dim, 3)
    conv("conv_before_upsample.0", num_feat, embed_dim, 3)
    conv("upsample.0", 4 * num_feat, num_feat, 3)
    conv("upsample.2", 4 * num_feat, num_feat, 3)
    conv("conv_last", 3, num_feat, 3)
    return s


def shift_mask(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """SwinTransformerBlock.calculate_mask, drct_arch.py:353-374: [nW, ws*ws, ws*ws] of 0 / -100."""
    img = torch.zeros(1, H, W, 1)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = _partition(img, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def synth_state_dict(shapes: Dict[str, Tuple[Tuple[int, ...], str]], seed: int = 0, window: Optional[int] = None,
                     img_size: int = 64) -> Dict[str, torch.Tensor]:
    """Deterministic non-degenerate weights for every float entry (one CPU generator per tensor, seeded from the
    tensor's name), exact buffers for the index / mask entries."""
    sd = {}
    for name, (shape, kind) in shapes.items():
        if kind == "index":
            sd[name] = relative_position_index(int(math.isqrt(shape[0])))
            continue
        if kind == "mask":
            ws = int(math.isqrt(shape[1]))
            sd[name] = shift_mask(img_size, img_size, ws, ws // 2)
            continue
        g = torch.Generator().manual_seed(int(hashlib.sha256(f"{seed}:{name}".encode()).hexdigest()[:8], 16))
        t = torch.randn(*shape, generator=g)
        if name.endswith("relative_position_bias_table"):
            t = 0.5 * t
        elif ".norm" in name or name.startswith("norm.") or name.startswith("patch_embed.norm"):
            t = (1.0 + 0.2 * t) if name.endswith(".weight") else 0.1 * t
        elif name.endswith(".bias"):
            t = 0.05 * t
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = t * (1.0 / math.sqrt(fan_in))
        sd[name] = t
    return sd


# ---------------------------------------------------------------------------------------------------
# forward
# ---------------------------------------------------------------------------------------------------
def _partition(x: torch.Tensor, ws: int) -> torch.Tensor:
    """window_partition, drct_arch.py:97-108."""
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def _reverse(w: torch.Tensor, ws: int, H: int, W: int) -> torch.Tensor:
    """window_reverse, drct_arch.py:111-124."""
    B = int(w.shape[0] / (H * W / ws / ws))
    x = w.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def window_attention(sd, p: str, xw: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """WindowAttention.forward, drct_arch.py:175-206: xw [nW*B, N, C]."""
    B_, N, C = xw.shape
    table = sd[p + ".relative_position_bias_table"]
    heads = table.shape[1]
    qkv = F.linear(xw, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"]).reshape(B_, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * ((C // heads) ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = table[sd[p + ".relative_position_index"].view(-1)].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(out, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def swin_block(sd, p: str, x: torch.Tensor, H: int, W: int, shifted: bool) -> torch.Tensor:
    """SwinTransformerBlock.forward, drct_arch.py:376-416 (tokens [B, H*W, C])."""
    B, L, C = x.shape
    ws = int(math.isqrt(sd[p + ".attn.relative_position_index"].shape[0]))
    shift = ws // 2 if shifted else 0
    h = _ln(sd, p + ".norm1", x).view(B, H, W, C)
    if shift:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
    mask = shift_mask(H, W, ws, shift).to(x.device) if shift else None
    a = window_attention(sd, p + ".attn", _partition(h, ws).view(-1, ws * ws, C), mask)
    h = _reverse(a.view(-1, ws, ws, C), ws, H, W)
    if shift:
        h = torch.roll(h, shifts=(shift, shift), dims=(1, 2))
    x = x + h.view(B, H * W, C)
    m = F.linear(_ln(sd, p + ".norm2", x), sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])
    m = F.linear(F.gelu(m), sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
    return x + m


def rdg(sd, p: str, x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """RDG.forward, drct_arch.py:292-300: dense growth through five swin blocks and 1x1 adjust convs (LeakyReLU 0.2)."""
    B = x.shape[0]
    feats = [x]
    for j in range(5):
        t = swin_block(sd, f"{p}.swin{j + 1}", torch.cat(feats, -1), H, W, shifted=(j % 2 == 1))
        img = t.transpose(1, 2).reshape(B, -1, H, W)
        img = F.conv2d(img, sd[f"{p}.adjust{j + 1}.weight"], sd[f"{p}.adjust{j + 1}.bias"])
        if j < 4:
            img = F.leaky_relu(img, 0.2)
        feats.append(img.flatten(2).transpose(1, 2))
    return feats[5] * 0.2 + x


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, img_range: float = 1.0, return_feature: bool = False):
    """DRCT.forward (pixelshuffle, '1conv'), drct_arch.py:761-789.  H, W must be multiples of the window (callers pad,
    scripts/extract_test_tta_cache.py).  ``return_feature``: also the ``conv_after_body`` output (the cached feature)."""
    n_rdg = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
    mean = torch.tensor(RGB_MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    x = (x - mean) * img_range
    x = F.conv2d(x, sd["conv_first.weight"], sd["conv_first.bias"], padding=1)
    B, C, H, W = x.shape
    t = _ln(sd, "patch_embed.norm", x.flatten(2).transpose(1, 2))
    for i in range(n_rdg):
        t = rdg(sd, f"layers.{i}", t, H, W)
    t = _ln(sd, "norm", t).transpose(1, 2).reshape(B, C, H, W)
    feat = F.conv2d(t, sd["conv_after_body.weight"], sd["conv_after_body.bias"], padding=1)
    y = feat + x
    y = F.leaky_relu(F.conv2d(y, sd["conv_before_upsample.0.weight"], sd["conv_before_upsample.0.bias"], padding=1), 0.01)
    y = F.pixel_shuffle(F.conv2d(y, sd["upsample.0.weight"], sd["upsample.0.bias"], padding=1), 2)
    y = F.pixel_shuffle(F.conv2d(y, sd["upsample.2.weight"], sd["upsample.2.bias"], padding=1), 2)
    y = F.conv2d(y, sd["conv_last.weight"], sd["conv_last.bias"], padding=1)
    y = y / img_range + mean
    return (y, feat) if return_feature else y


def flops_per_lr_pixel(embed_dim=180, n_rdg=12, window=16, num_heads=6, gc=32, mlp_ratio=2, num_feat=64) -> float:
    """2 x MAC per LR pixel of the forward (linears, attention products, convs) -- the roofline unit of N1."""
    f = 2 * 9 * 3 * embed_dim
    N = window * window
    for (d, h, hid, _), j in zip(swin_dims(embed_dim, num_heads, gc, mlp_ratio), range(5)):
        blk = 2 * (3 * d * d + d * d + 2 * d * hid) + 2 * 2 * N * d
        blk += 2 * d * (gc if j < 4 else embed_dim)
        f += n_rdg * blk
    f += 2 * 9 * embed_dim * embed_dim + 2 * 9 * embed_dim * num_feat
    f += 2 * 9 * num_feat * 4 * num_feat + 4 * 2 * 9 * num_feat * 4 * num_feat + 16 * 2 * 9 * num_feat * 3
    return float(f)
