"""torch.library registration of the C-ABI kernels (isr_b200.torch_ops): registration, schemas and fake kernels on the CPU;
values and torch.library.opcheck on the GPU."""
import pytest
import torch

import isr_b200
from isr_b200 import torch_ops


def test_ops_are_registered_with_cuda_only_kernels():
    for n in torch_ops.OPS:
        assert hasattr(torch.ops.ffsr, n), n
    assert str(torch.ops.ffsr.fusion_forward.default._schema) == \
        "ffsr::fusion_forward(Tensor lr, Tensor[] expert_imgs, Tensor[] expert_feats, Tensor[] state, str precision=\"bf16\") -> Tensor"
    with pytest.raises(NotImplementedError):                      # no CPU backend: the dispatcher says so, nothing falls back
        torch.ops.ffsr.layernorm(torch.randn(4, 64), torch.ones(64), torch.zeros(64))


def test_fake_kernels_give_output_shapes_without_a_gpu():
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        lr = torch.empty(2, 3, 8, 12, device="cuda")
        out = torch.ops.ffsr.fusion_forward(lr, [torch.empty(2, 3, 32, 48, device="cuda")] * 4, [], [], "bf16")
        assert tuple(out.shape) == (2, 3, 32, 48) and out.dtype == torch.float32
        o3, at = torch.ops.ffsr.edge_refiner(torch.empty(1, 20, 30, 8, device="cuda", dtype=torch.bfloat16),
                                             torch.empty(23552, device="cuda", dtype=torch.bfloat16), torch.empty(320, device="cuda"))
        assert tuple(o3.shape) == (1, 20, 30, 32) and tuple(at.shape) == (1, 20, 30)


@pytest.mark.gpu
def test_ops_match_their_torch_definitions_on_the_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import torch.nn.functional as F
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 7, 128, generator=g).to(dev)
    w, b = torch.randn(128, generator=g).to(dev), torch.randn(128, generator=g).to(dev)
    y = torch.ops.ffsr.layernorm(x, w, b)
    assert (y - F.layer_norm(x, (128,), w, b, 1e-5)).abs().max().item() < 2e-5
    torch.library.opcheck(torch.ops.ffsr.layernorm.default, (x, w, b), test_utils=("test_schema", "test_faketensor"))
    # token attention vs softmax(q k^T / 4) v per pixel and head
    B, T, HW, E = 2, 4, 37, 128
    qkv = torch.randn(B, T, HW, 3 * E, generator=g).to(dev)
    ctx = torch.ops.ffsr.token_attention(qkv)
    q, k, v = [t.permute(0, 2, 1, 3).reshape(B * HW, T, E // 16, 16).transpose(1, 2) for t in qkv.split(E, dim=-1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 4.0, -1) @ v).transpose(1, 2).reshape(B, HW, T, E).permute(0, 2, 1, 3)
    assert (ctx - ref).abs().max().item() < 2e-5
    # whole forward as one op == the module
    torch.manual_seed(0)
    m = isr_b200.CompleteEnhancedFusionSR(None).eval().to(dev)
    m.precision = "fp32"
    from oracle import fusion_oracle as O
    lr, imgs, fts, _ = O.synthetic_inputs(1, 16, 16)
    lr, imgs, fts = lr.to(dev), {k: v.to(dev) for k, v in imgs.items()}, {k: v.to(dev) for k, v in fts.items()}
    want = m.forward_with_precomputed(lr, imgs, fts)
    got = torch.ops.ffsr.fusion_forward(lr, [imgs[k] for k in isr_b200.EXPERT_ORDER], [fts[k] for k in isr_b200.EXPERT_ORDER],
                                        list(m.state_dict().values()), "fp32")
    assert torch.equal(got, want)
