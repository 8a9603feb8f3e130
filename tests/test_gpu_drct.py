"""DRCT-L window attention: the (shifted-)window attention core of every SwinTransformerBlock against a
torch restatement of drct_arch.py (itself checked against the oracle block on the CPU); argument validation on the CPU."""
import ctypes as C
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import isr_b200  # noqa: E402,F401
from isr_b200 import _cabi as K  # noqa: E402
from oracle import drct_oracle as DO  # noqa: E402


def expected_attention(qkv: torch.Tensor, heads: int, ws: int, shift: int, table: torch.Tensor) -> torch.Tensor:
    """drct_arch.py:385-412 around WindowAttention.forward :175-206 without the two Linears (torch ops, fp32)."""
    B, H, W, C3 = qkv.shape
    Cc = C3 // 3
    dh, N = Cc // heads, ws * ws
    x = torch.roll(qkv, shifts=(-shift, -shift), dims=(1, 2)) if shift else qkv
    xw = DO._partition(x, ws).view(-1, N, 3, heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = xw[0] * dh ** -0.5, xw[1], xw[2]
    attn = q @ k.transpose(-2, -1)
    bias = table[DO.relative_position_index(ws).view(-1)].view(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if shift:
        mask = DO.shift_mask(H, W, ws, shift)
        nW = mask.shape[0]
        attn = (attn.view(-1, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    out = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, ws, ws, Cc)
    out = DO._reverse(out, ws, H, W)
    return torch.roll(out, shifts=(shift, shift), dims=(1, 2)) if shift else out


def test_expected_attention_is_the_oracle_block_without_linears():
    """The torch restatement used as the checker equals oracle.swin_block's attention core (identity qkv / proj)."""
    torch.manual_seed(0)
    Cc, heads, ws = 12, 2, 4
    H, W = 8, 12
    x = torch.randn(1, H, W, Cc)
    Wqkv = torch.randn(3 * Cc, Cc) / Cc ** 0.5
    table = 0.5 * torch.randn((2 * ws - 1) ** 2, heads)
    for shift in (0, 2):
        sd = {"a.qkv.weight": Wqkv, "a.qkv.bias": torch.zeros(3 * Cc), "a.proj.weight": torch.eye(Cc), "a.proj.bias": torch.zeros(Cc),
              "a.relative_position_bias_table": table, "a.relative_position_index": DO.relative_position_index(ws)}
        h = torch.roll(x, (-shift, -shift), (1, 2)) if shift else x
        mask = DO.shift_mask(H, W, ws, shift) if shift else None
        want = DO._reverse(DO.window_attention(sd, "a", DO._partition(h, ws).view(-1, ws * ws, Cc), mask).view(-1, ws, ws, Cc), ws, H, W)
        want = torch.roll(want, (shift, shift), (1, 2)) if shift else want
        got = expected_attention(torch.nn.functional.linear(x, Wqkv), heads, ws, shift, table)
        assert float((got - want).abs().max()) < 1e-5


def test_window_attention_rejects_bad_arguments():
    lib = K.load()
    f = lib.ffsr_window_attention
    ok = (4096, 1, 16, 16, 180, 6, 8, 4, 4096, 4096, K.DT_F32, None)
    bad = [
        (0,) + ok[1:],                                        # null qkv
        ok[:2] + (15,) + ok[3:],                              # H not a multiple of the window
        ok[:5] + (7,) + ok[6:],                               # C % heads != 0
        ok[:6] + (17,) + ok[7:],                              # window^2 > 256
        ok[:7] + (8,) + ok[8:],                               # shift >= window
        ok[:10] + (K.DT_F16,) + ok[11:],                      # unsupported dtype
    ]
    for a in bad:
        assert f(*a) == -1, a
    assert b"window_attention" in lib.ffsr_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("Cc,heads,ws,dtype", [(180, 6, 16, "fp32"), (212, 4, 16, "fp32"), (244, 2, 16, "bf16"), (276, 6, 8, "fp32"),
                                               (308, 4, 16, "bf16"), (60, 6, 8, "fp32"),
                                               # fp32 K + V of head dim 122 do not fit in shared memory: K staged, V read through L2
                                               (244, 2, 16, "fp32")])
def test_window_attention_matches_the_reference_block(Cc, heads, ws, dtype):
    dev = torch.device("cuda:0")
    lib = K.load()
    g = torch.Generator().manual_seed(Cc + ws)
    B, H, W = 2, 2 * ws, 3 * ws
    qkv = torch.randn(B, H, W, 3 * Cc, generator=g)
    table = 0.5 * torch.randn((2 * ws - 1) ** 2, heads, generator=g)
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    qd = qkv.to(dev, tdt).contiguous()
    td = table.to(dev)
    for shift in (0, ws // 2):
        out = torch.empty(B, H, W, Cc, device=dev, dtype=tdt)
        rc = lib.ffsr_window_attention(qd.data_ptr(), B, H, W, Cc, heads, ws, shift, td.data_ptr(), out.data_ptr(),
                                       K.DT_BF16 if dtype == "bf16" else K.DT_F32, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        K.check(rc, "window_attention")
        want = expected_attention(qd.float().cpu(), heads, ws, shift, table)
        err = float((out.float().cpu() - want).abs().max())
        assert err <= (2e-2 if dtype == "bf16" else 2e-5), (Cc, heads, ws, shift, err)


@pytest.mark.gpu
def test_drct_forward_matches_oracle():
    """isr_b200.drct.DRCT.forward (fp32, C ABI kernels) against oracle/drct_oracle.py on the reduced configuration of the
    golden (2 RDGs, window 8, all DRCT-L channel counts), synthesised weights, a 16x24 input (shifted windows, 2x3 grid)."""
    import json
    import numpy as np
    from isr_b200 import drct as D
    g = np.load(os.path.join(ROOT, "tests", "golden", "drct_small.npz"))
    cfg = json.loads(str(g["cfg"]))
    m = D.DRCT(img_size=cfg["img_size"], window_size=cfg["window"], embed_dim=cfg["embed_dim"], depths=[6] * cfg["n_rdg"],
               num_heads=[cfg["num_heads"]] * cfg["n_rdg"], mlp_ratio=cfg["mlp_ratio"])
    sd = DO.synth_state_dict(DO.state_shapes(**cfg), seed=int(g["seed"]), img_size=cfg["img_size"])
    m.load_state_dict(sd, strict=True)
    dev = torch.device("cuda:0")
    m.to(dev).eval()
    x = torch.from_numpy(g["x"])
    y = m(x.to(dev))
    torch.cuda.synchronize()
    assert float((y.cpu() - torch.from_numpy(g["y"])).abs().max()) <= 1e-4           # the reference class's own output
    with torch.no_grad():
        want, feat = DO.forward(sd, x, return_feature=True)
    assert float((y.cpu() - want).abs().max()) <= 1e-4
    assert float((m.last_feature.cpu() - feat).abs().max()) <= 1e-4




@pytest.mark.gpu
@pytest.mark.parametrize("Cc,heads,ws", [(180, 6, 16), (212, 4, 16), (244, 2, 16), (276, 6, 16), (308, 4, 16), (308, 4, 8)])
def test_pitched_window_attention_bf16(Cc, heads, ws):
    dev = torch.device("cuda:0")
    lib = K.load()
    g = torch.Generator().manual_seed(Cc)
    B, H, W = 1, 2 * ws, 2 * ws
    qp, op = (3 * Cc + 7) // 8 * 8, (Cc + 7) // 8 * 8
    qkv = torch.randn(B, H, W, 3 * Cc, generator=g)
    table = 0.5 * torch.randn((2 * ws - 1) ** 2, heads, generator=g)
    qd = torch.zeros(B, H, W, qp, device=dev, dtype=torch.bfloat16)
    qd[..., :3 * Cc] = qkv.to(dev)
    td = table.to(dev)
    for shift in (0, ws // 2):
        out = torch.full((B, H, W, op), 7.0, device=dev, dtype=torch.bfloat16)
        K.check(lib.ffsr_window_attention_pitched(qd.data_ptr(), qp, B, H, W, Cc, heads, ws, shift, td.data_ptr(), out.data_ptr(), op,
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "window_attention_pitched")
        want = expected_attention(qd[..., :3 * Cc].float().cpu(), heads, ws, shift, table)
        assert float((out[..., :Cc].float().cpu() - want).abs().max()) <= 2e-2
        assert bool((out[..., Cc:] == 7.0).all())                      # padding channels are not touched


@pytest.mark.gpu
def test_drct_forward_bf16_mode():
    import json
    import numpy as np
    from isr_b200 import drct as D
    g = np.load(os.path.join(ROOT, "tests", "golden", "drct_small.npz"))
    cfg = json.loads(str(g["cfg"]))
    m = D.DRCT(img_size=cfg["img_size"], window_size=cfg["window"], embed_dim=cfg["embed_dim"], depths=[6] * cfg["n_rdg"],
               num_heads=[cfg["num_heads"]] * cfg["n_rdg"], mlp_ratio=cfg["mlp_ratio"])
    m.load_state_dict(DO.synth_state_dict(DO.state_shapes(**cfg), seed=int(g["seed"]), img_size=cfg["img_size"]), strict=True)
    dev = torch.device("cuda:0")
    m.to(dev).eval()
    m.precision = "bf16"
    y = m(torch.from_numpy(g["x"]).to(dev)).cpu()
    want = torch.from_numpy(g["y"])
    rel = float((y - want).norm() / want.norm())
    assert rel <= 2e-2, rel
