"""``torch.library`` registration of the C-ABI kernels: the dispatcher-visible face of ``libffsr_b200.so``.

The library itself has a plain C ABI (``include/ffsr_b200.h``: raw device pointers, sizes and a ``cudaStream_t`` -- no torch
types, so any host language can bind it) and is loaded with ``ctypes`` (``_cabi.py``).  This module registers the
tensor-in / tensor-out entry points as custom operators in the ``ffsr`` namespace,

    torch.ops.ffsr.fusion_forward, frequency_bands, layernorm, token_attention, lka_depthwise, edge_refiner,
    modulate_hr, fused_losses

so that they are ordinary PyTorch ops: visible to the dispatcher, usable under ``torch.no_grad`` / ``torch.compile`` tracing
(each has a fake / meta implementation that only computes output shapes), listed by ``torch.library`` introspection and
checked by ``torch.library.opcheck``.  Only a CUDA implementation exists: on CPU tensors the dispatcher reports the missing
kernel (there is no CPU fallback, by design).  ``CompleteEnhancedFusionSR`` itself keeps calling the C ABI directly; the
whole-model op ``ffsr::fusion_forward`` wraps it for callers that want one dispatcher-visible node per forward.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from . import _cabi as K

_NS = "ffsr"
EXPERT_ORDER = ("drct", "grl", "nafnet", "mamba")


def _S(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ffsr ops (sm_100a build) need CUDA tensors: there is no CPU path")


# ---------------------------------------------------------------------------------------------------
# nn.LayerNorm over the last dim (E in {64,128,256}) of dense rows        large_kernel_attention.py:385-394
# ---------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{_NS}::layernorm", mutates_args=(), device_types="cuda")
def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, out_bf16: bool = False) -> torch.Tensor:
    _cuda(x, weight, bias)
    E = x.shape[-1]
    xf = x.detach().float().contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    with torch.cuda.device(x.device):
        K.check(K.load().ffsr_layernorm(xf.data_ptr(), xf.numel() // E, E, weight.detach().float().contiguous().data_ptr(),
                                        bias.detach().float().contiguous().data_ptr(), y.data_ptr(), int(out_bf16), _S(x)), "layernorm")
    return y


@layernorm.register_fake
def _(x, weight, bias, out_bf16=False):
    return torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)


# ---------------------------------------------------------------------------------------------------
# softmax(q k^T / 4) v over the T tokens of each LR pixel, head_dim 16; qkv [B][T][HW][3E] -> ctx [B][T][HW][E]
# ---------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{_NS}::token_attention", mutates_args=(), device_types="cuda")
def token_attention(qkv: torch.Tensor) -> torch.Tensor:
    _cuda(qkv)
    if qkv.dim() != 4 or qkv.shape[-1] % 48 or qkv.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("token_attention: qkv must be fp32 / bf16 [B, T, HW, 3E] with E a multiple of 16")
    B, T, HW, E3 = qkv.shape
    q = qkv.detach().contiguous()
    ctx = torch.empty(B, T, HW, E3 // 3, device=qkv.device, dtype=qkv.dtype)
    with torch.cuda.device(qkv.device):
        K.check(K.load().ffsr_token_attention(q.data_ptr(), B, T, HW, E3 // 3, ctx.data_ptr(), int(qkv.dtype == torch.bfloat16), _S(qkv)),
                "token_attention")
    return ctx


@token_attention.register_fake
def _(qkv):
    B, T, HW, E3 = qkv.shape
    return torch.empty(B, T, HW, E3 // 3, device=qkv.device, dtype=qkv.dtype)


# ---------------------------------------------------------------------------------------------------
# LKA depthwise chain: BN-fold -> dw5x5 -> dw1x21 -> dw21x1 on channels-last [N][H][W][C]   large_kernel_attention.py:96-105
# ---------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{_NS}::lka_depthwise", mutates_args=(), device_types="cuda")
def lka_depthwise(x: torch.Tensor, bn_scale: torch.Tensor, bn_shift: torch.Tensor, w5: torch.Tensor, wh: torch.Tensor,
                  wv: torch.Tensor, out_bf16: bool = False) -> torch.Tensor:
    _cuda(x, bn_scale, bn_shift, w5, wh, wv)
    if x.dim() != 4 or x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("lka_depthwise: x must be fp32 / bf16 channels-last [N, H, W, C]")
    N, H, W, Cc = x.shape
    xc = x.detach().contiguous()
    f = lambda t, n: t.detach().float().reshape(Cc, n).contiguous()           # noqa: E731
    t1 = torch.empty(N, H, W, Cc, device=x.device)
    t2 = torch.empty_like(t1)
    out = torch.empty(N, H, W, Cc, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    k, d = bn_scale.detach().float().contiguous(), bn_shift.detach().float().contiguous()
    a5, ah, av = f(w5, 25), f(wh, 21), f(wv, 21)
    with torch.cuda.device(x.device):
        K.check(K.load().ffsr_lka_depthwise_in(xc.data_ptr(), K.DT_BF16 if x.dtype == torch.bfloat16 else K.DT_F32, N, H, W, Cc,
                                               k.data_ptr(), d.data_ptr(), a5.data_ptr(), ah.data_ptr(), av.data_ptr(), t1.data_ptr(),
                                               t2.data_ptr(), out.data_ptr(), K.DT_BF16 if out_bf16 else K.DT_F32, _S(x)), "lka_depthwise")
    return out


@lka_depthwise.register_fake
def _(x, bn_scale, bn_shift, w5, wh, wv, out_bf16=False):
    return torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)


# ---------------------------------------------------------------------------------------------------
# one pyramid level of the edge refiner as ONE tile-resident tcgen05 kernel          edge_enhancement.py:69-118
# x: bf16 [N][H][W][8] (3 Laplacian channels); returns (o3 bf16 [N][H][W][32], attention fp32 [N][H][W])
# ---------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{_NS}::edge_refiner", mutates_args=(), device_types="cuda")
def edge_refiner(x: torch.Tensor, weight_blob: torch.Tensor, param_blob: torch.Tensor) -> List[torch.Tensor]:
    _cuda(x, weight_blob, param_blob)
    lib = K.load()
    if x.dim() != 4 or x.shape[-1] != 8 or x.dtype != torch.bfloat16:
        raise ValueError("edge_refiner: x must be bf16 channels-last [N, H, W, 8]")
    if weight_blob.numel() * weight_blob.element_size() != lib.ffsr_edge_chain_weight_bytes() or \
            param_blob.numel() != lib.ffsr_edge_chain_param_floats() or param_blob.dtype != torch.float32:
        raise ValueError("edge_refiner: blobs must come from isr_b200.pipeline.pack_edge_chain")
    N, H, W, _ = x.shape
    xc = x.detach().contiguous()
    o3 = torch.empty(N, H, W, 32, device=x.device, dtype=torch.bfloat16)
    at = torch.empty(N, H, W, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        K.check(lib.ffsr_edge_refiner_chain(xc.data_ptr(), N, H, W, weight_blob.contiguous().data_ptr(), param_blob.contiguous().data_ptr(),
                                            None, 0, o3.data_ptr(), H * W * 32, W * 32, 32, at.data_ptr(), 1, _S(x)), "edge_refiner_chain")
    return [o3, at]


@edge_refiner.register_fake
def _(x, weight_blob, param_blob):
    N, H, W, _ = x.shape
    return [torch.empty(N, H, W, 32, device=x.device, dtype=torch.bfloat16), torch.empty(N, H, W, device=x.device, dtype=torch.float32)]


# ---------------------------------------------------------------------------------------------------
# the whole cached-fusion forward as ONE dispatcher-visible op.  ``state`` = the module's state_dict values in key order
# (fusion.py: 226 tensors); expert tensors in the order drct, grl, nafnet, mamba.
# ---------------------------------------------------------------------------------------------------
_MODELS = {}


def _model_for(state: Sequence[torch.Tensor], precision: str):
    from .fusion import CompleteEnhancedFusionSR
    dev = state[0].device
    key = (str(dev), precision)
    m = _MODELS.get(key)
    if m is None:
        m = CompleteEnhancedFusionSR(None).eval().to(dev)
        m.precision = precision
        _MODELS[key] = m
    keys = list(m.state_dict().keys())
    if len(keys) != len(state):
        raise ValueError(f"fusion_forward: expected the {len(keys)} state_dict tensors in key order, got {len(state)}")
    cur = m.state_dict()
    if any(cur[k].data_ptr() != v.data_ptr() and not torch.equal(cur[k], v) for k, v in zip(keys, state)):
        m.load_state_dict(dict(zip(keys, state)))
    return m


@torch.library.custom_op(f"{_NS}::fusion_forward", mutates_args=(), device_types="cuda")
def fusion_forward(lr: torch.Tensor, expert_imgs: List[torch.Tensor], expert_feats: List[torch.Tensor], state: List[torch.Tensor],
                   precision: str = "bf16") -> torch.Tensor:
    _cuda(lr, *expert_imgs, *expert_feats)
    if len(expert_imgs) != 4 or len(expert_feats) not in (0, 4):
        raise ValueError("fusion_forward: 4 expert images (drct, grl, nafnet, mamba) and 0 or 4 feature maps")
    m = _model_for(state, precision)
    with torch.no_grad():
        return m.forward_with_precomputed(lr, dict(zip(EXPERT_ORDER, expert_imgs)),
                                          dict(zip(EXPERT_ORDER, expert_feats)) if expert_feats else None)


@fusion_forward.register_fake
def _(lr, expert_imgs, expert_feats, state, precision="bf16"):
    B, _, H, W = lr.shape
    return torch.empty(B, 3, 4 * H, 4 * W, device=lr.device, dtype=torch.float32)


OPS = ("layernorm", "token_attention", "lka_depthwise", "edge_refiner", "fusion_forward")
