"""GPU parity of the train-mode forward/backward (through the C ABI) against autograd over the
CPU oracle.  Run on the B200 box: ``pytest tests -m gpu``.

Tolerances: forward max-abs <= 1e-4; all 198 parameter gradients rel-L2 <= 1e-4
(fp32 path, attention dropout 0); BatchNorm running statistics <= 1e-6, counters exact.
"""
import pytest
import torch
import torch.nn.functional as F

import isr_b200
from isr_b200 import _cabi as K
from isr_b200 import training as T
from oracle import fusion_oracle as O
from oracle.perturb import perturb_state_dict

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


# ------------------------------------------------------------------------------------------
# primitives against torch autograd on the CPU
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cin,cout,ks,layout", [
    (3, 32, 3, "nchw"), (12, 64, 3, "cl"), (76, 64, 3, "cl"), (128, 128, 3, "cl"), (128, 3, 3, "cl"),
    (32, 1, 3, "cl"), (180, 128, 1, "nchw"), (128, 384, 1, "cl"), (256, 128, 1, "cl"), (16, 4, 1, "cl"),
    (6, 16, 3, "cl"), (64, 3, 1, "cl"),
])
def test_conv2d_forward_backward(cin, cout, ks, layout):
    dev = _cuda()
    g = torch.Generator().manual_seed(cin * 131 + cout * 7 + ks)
    N, H, W = 2, 19, 37
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, ks, ks, generator=g) / (cin * ks * ks) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    gy = torch.randn(N, cout, H, W, generator=g)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    F.conv2d(xr, wr, br, padding=ks // 2).backward(gy)
    xd = x.to(dev)
    if layout == "cl":
        xd = xd.contiguous(memory_format=torch.channels_last)
    xd.requires_grad_()
    wd, bd = w.to(dev).requires_grad_(), b.to(dev).requires_grad_()
    y = T.conv2d(xd, wd, bd)
    y.backward(gy.to(dev))
    torch.cuda.synchronize()
    assert _rel(y, F.conv2d(x, w, b, padding=ks // 2)) < 2e-6
    assert _rel(xd.grad, xr.grad) < 2e-6
    assert _rel(wd.grad, wr.grad) < 5e-6
    assert _rel(bd.grad, br.grad) < 5e-6


@pytest.mark.parametrize("kind,fn", [(K.ACT_GELU, F.gelu), (K.ACT_RELU, F.relu), (K.ACT_SIGMOID, torch.sigmoid)])
def test_activation_forward_backward(kind, fn):
    dev = _cuda()
    g = torch.Generator().manual_seed(kind)
    x = torch.randn(2, 5, 13, 17, generator=g) * 2
    gy = torch.randn(2, 5, 13, 17, generator=g)
    xr = x.clone().requires_grad_()
    fn(xr).backward(gy)
    xd = x.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    y = T._Act.apply(xd, kind)
    y.backward(gy.to(dev))
    assert (y.cpu() - fn(x)).abs().max() < 2e-6
    assert (xd.grad.cpu() - xr.grad).abs().max() < 2e-6


@pytest.mark.parametrize("G", [1, 3])
def test_batchnorm_train_forward_backward(G):
    dev = _cuda()
    g = torch.Generator().manual_seed(5 + G)
    B, Cc, H, W = 2, 64, 9, 11
    x = torch.randn(G * B, Cc, H, W, generator=g) * 1.7 + 0.3
    gy = torch.randn(G * B, Cc, H, W, generator=g)
    bn = torch.nn.BatchNorm2d(Cc)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(Cc, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(Cc, generator=g) * 0.1)
    ref = torch.nn.BatchNorm2d(Cc)
    ref.load_state_dict(bn.state_dict())
    ref.train()
    xr = x.clone().requires_grad_()
    ys = [ref(xr[i * B:(i + 1) * B]) for i in range(G)]         # G sequential calls, as the LKABlock loop does
    torch.cat(ys).backward(gy)
    bn = bn.to(dev)
    sink = []
    xd = x.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    y = T.batchnorm_train(xd, bn, G, sink)
    y.backward(gy.to(dev))
    T.fold_running_stats(sink)
    assert (y.cpu() - torch.cat(ys)).abs().max() < 5e-6
    assert _rel(xd.grad, xr.grad) < 5e-6
    assert _rel(bn.weight.grad, ref.weight.grad) < 5e-6
    assert _rel(bn.bias.grad, ref.bias.grad) < 5e-6
    assert (bn.running_mean.cpu() - ref.running_mean).abs().max() < 1e-6
    assert (bn.running_var.cpu() - ref.running_var).abs().max() < 1e-6
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == G


def test_layernorm_forward_backward():
    dev = _cuda()
    g = torch.Generator().manual_seed(11)
    N, E, H, W = 3, 128, 7, 9
    x = torch.randn(N, E, H, W, generator=g)
    gy = torch.randn(N, E, H, W, generator=g)
    ln = torch.nn.LayerNorm(E)
    with torch.no_grad():
        ln.weight.copy_(torch.rand(E, generator=g) + 0.5)
        ln.bias.copy_(torch.randn(E, generator=g) * 0.1)
    xr = x.clone().requires_grad_()
    yr = ln(xr.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    yr.backward(gy)
    wg, bg = ln.weight.grad.clone(), ln.bias.grad.clone()
    ln.zero_grad()
    ln = ln.to(dev)
    xd = x.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    y = T.layernorm(xd, ln)
    y.backward(gy.to(dev))
    assert (y.cpu() - yr).abs().max() < 5e-6
    assert _rel(xd.grad, xr.grad) < 5e-6
    assert _rel(ln.weight.grad, wg) < 5e-6
    assert _rel(ln.bias.grad, bg) < 5e-6


@pytest.mark.parametrize("Tn,E,heads", [(4, 128, 8), (9, 64, 4)])
def test_token_attention_forward_backward(Tn, E, heads):
    dev = _cuda()
    g = torch.Generator().manual_seed(Tn)
    B, H, W = 2, 5, 7
    mha = torch.nn.MultiheadAttention(E, heads, batch_first=True, dropout=0.0)
    x = torch.randn(B, Tn, H, W, E, generator=g)                  # token-major memory
    gy = torch.randn(B, Tn, H, W, E, generator=g)
    xr = x.clone().requires_grad_()
    seq = xr.permute(0, 2, 3, 1, 4).reshape(B * H * W, Tn, E)
    yr = mha(seq, seq, seq, need_weights=False)[0].reshape(B, H, W, Tn, E).permute(0, 3, 1, 2, 4)
    yr.backward(gy)
    ref_g = {n: p.grad.clone() for n, p in mha.named_parameters()}
    mha.zero_grad()
    mha = mha.to(dev)
    xd = x.to(dev).reshape(B * Tn, H, W, E).permute(0, 3, 1, 2).requires_grad_()
    y = T.mha_tokens(xd, mha, B, Tn, training=True)
    y.backward(gy.to(dev).reshape(B * Tn, H, W, E).permute(0, 3, 1, 2))
    assert (y.permute(0, 2, 3, 1).reshape(B, Tn, H, W, E).cpu() - yr).abs().max() < 5e-6
    assert _rel(xd.grad.permute(0, 2, 3, 1).reshape(B, Tn, H, W, E), xr.grad) < 1e-5
    for n, p in mha.named_parameters():
        assert _rel(p.grad, ref_g[n]) < 1e-5, n


def test_token_attention_dropout_statistics():
    """Dropout on the attention probabilities: keep-rate ~ 1-p, rescaled by 1/(1-p), and the
    backward regenerates the same mask (gradient of sum(ctx) w.r.t. v counts kept probabilities)."""
    dev = _cuda()
    B, Tn, E, H, W = 2, 4, 128, 16, 16
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B * Tn, 3 * E, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    qkv[:, :2 * E] = 0                                              # uniform attention: p = 1/T
    qkv[:, 2 * E:] = 1.0                                            # v = 1  -> ctx = sum of kept p / (1-p)
    qkv.requires_grad_()
    ctx = T._TokenAttention.apply(qkv, B, Tn, 0.1, 12345)
    mean = ctx.mean().item()
    assert abs(mean - 1.0) < 0.01, mean                             # unbiased
    vals = torch.unique((ctx.detach() * 0.9 * Tn).round())
    assert set(vals.tolist()) <= set(range(Tn + 1))                 # k kept of T, each worth 1/(T*0.9)
    ctx.sum().backward()
    gv = qkv.grad[:, 2 * E:]
    assert abs(gv.mean().item() - 1.0) < 0.01
    ctx2 = T._TokenAttention.apply(qkv.detach(), B, Tn, 0.1, 12345)
    assert torch.equal(ctx2, ctx.detach())                          # same seed, same mask


@pytest.mark.parametrize("kind,kh,kw", [(0, 5, 5), (1, 1, 21), (2, 21, 1)])
def test_depthwise_stage_forward_backward(kind, kh, kw):
    dev = _cuda()
    g = torch.Generator().manual_seed(kind)
    N, Cc, H, W = 2, 64, 23, 19
    x = torch.randn(N, Cc, H, W, generator=g)
    w = torch.randn(Cc, 1, kh, kw, generator=g) * 0.2
    gy = torch.randn(N, Cc, H, W, generator=g)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    yr = F.conv2d(xr, wr, padding=(kh // 2, kw // 2), groups=Cc)
    yr.backward(gy)
    xd = x.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    wd = w.to(dev).requires_grad_()
    y = T._DwStage.apply(xd, wd, kind)
    y.backward(gy.to(dev))
    assert (y.cpu() - yr).abs().max() < 1e-5
    assert _rel(xd.grad, xr.grad) < 5e-6
    assert _rel(wd.grad, wr.grad) < 5e-6


# ------------------------------------------------------------------------------------------
# whole train-mode forward + backward against autograd over the oracle
# ------------------------------------------------------------------------------------------
def _train_model(seed=0):
    torch.manual_seed(seed)
    m = isr_b200.CompleteEnhancedFusionSR(None)
    m.load_state_dict(perturb_state_dict(m.state_dict(), seed=7), strict=True)
    m.cross_band.band_attention.dropout = 0.0
    m.collaborative.cross_attn.dropout = 0.0
    return m.train()


@pytest.mark.parametrize("B,H,W,with_feats", [(2, 12, 12, True), (1, 17, 23, True), (2, 9, 11, False)])
def test_train_forward_backward_matches_oracle_autograd(B, H, W, with_feats):
    dev = _cuda()
    m = _train_model()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    lr, imgs, fts, hr = O.synthetic_inputs(B, H, W, feats=with_feats)

    # oracle: fp32 autograd on the CPU
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() and k in dict(m.named_parameters()) else v.clone())
          for k, v in sd0.items()}
    upd = {}
    ref = O.run_pipeline(sd, lr, imgs, fts, training=True, bn_updates=upd)
    loss_ref = ((ref - hr) ** 2).mean() * 100.0
    loss_ref.backward()

    m.to(dev)
    out = m.forward_with_precomputed(lr.to(dev), {k: v.to(dev) for k, v in imgs.items()},
                                     {k: v.to(dev) for k, v in fts.items()} if fts else None)
    assert out.requires_grad and out.dtype == torch.float32 and tuple(out.shape) == (B, 3, 4 * H, 4 * W)
    loss = ((out - hr.to(dev)) ** 2).mean() * 100.0
    loss.backward()
    torch.cuda.synchronize()
    assert (out.detach().cpu() - ref.detach()).abs().max().item() <= 1e-4
    n_params, worst = 0, (0.0, "")
    for name, p in m.named_parameters():
        g_ref = sd[name].grad
        if not with_feats and name.startswith("collaborative."):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        assert p.grad is not None, f"{name}: no gradient"
        n_params += 1
        assert g_ref is not None, name
        denom = g_ref.double().norm().item()
        err = (p.grad.double().cpu() - g_ref.double()).norm().item()
        r = err / denom if denom > 1e-12 else err
        if r > worst[0]:
            worst = (r, name)
        assert r <= 1e-4, f"{name}: rel-L2 {r:.3e} (|g_ref| {denom:.3e})"
    if with_feats:
        assert n_params == 198
    # BatchNorm side effects
    after = m.state_dict()
    for k, v in upd.items():
        if k.endswith("num_batches_tracked"):
            assert int(after[k]) == int(v), k
        else:
            assert (after[k].cpu() - v).abs().max().item() <= 1e-6, k
    print(f"worst gradient rel-L2 {worst[0]:.2e} at {worst[1]}")


def test_train_step_with_dropout_runs_and_is_seeded():
    dev = _cuda()
    torch.manual_seed(0)
    m = isr_b200.CompleteEnhancedFusionSR(None).train().to(dev)
    lr, imgs, fts, hr = O.synthetic_inputs(2, 16, 16)
    args = (lr.to(dev), {k: v.to(dev) for k, v in imgs.items()}, {k: v.to(dev) for k, v in fts.items()})
    torch.manual_seed(5)
    a = m.forward_with_precomputed(*args)
    torch.manual_seed(5)
    b = m.forward_with_precomputed(*args)
    torch.manual_seed(6)
    c = m.forward_with_precomputed(*args)
    assert (a - b).abs().max().item() < 1e-6          # same seed -> same dropout masks (fp64 atomics aside)
    assert (a - c).abs().max().item() > 1e-6          # different seed -> different masks
    F.l1_loss(a.clamp(0, 1), hr.to(dev)).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    assert int(m.cross_band.lka_block.norm1.num_batches_tracked) == 27        # 3 forwards x 9 band calls
    assert int(m.collaborative.lka_global.norm1.num_batches_tracked) == 12    # 3 forwards x 4 expert calls
