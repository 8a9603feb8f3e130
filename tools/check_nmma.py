"""Bit-compare k_conv_tc outputs across issue configurations (FFSR_TC_NMMA / FFSR_TC_NACC / FFSR_TC_LEAN0 ...): the
math order per tile does not depend on them, so every output must be IDENTICAL to the single-issuer launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import isr_b200  # noqa: E402
from isr_b200 import _cabi as K  # noqa: E402
from isr_b200.pipeline import FusionEngine, nhwc, _pack_conv  # noqa: E402

dev = torch.device("cuda:0")
eng = FusionEngine(isr_b200.CompleteEnhancedFusionSR(None))
eng._stream = eng._get_stream(dev)
shapes = [(32, 32, 3, K.ACT_GELU, True), (64, 64, 3, K.ACT_GELU, True), (76, 64, 3, K.ACT_GELU, True), (3, 32, 3, K.ACT_GELU, True),
          (32, 8, 1, K.ACT_GELU, True), (8, 1, 3, K.ACT_SIGMOID, False), (16, 3, 3, K.ACT_SIGMOID, False), (128, 3, 3, K.ACT_NONE, False),
          (3, 128, 3, K.ACT_GELU, True), (128, 128, 1, K.ACT_NONE, False), (96, 32, 3, K.ACT_GELU, True), (32, 32, 3, K.ACT_NONE, False)]
configs = [{"FFSR_TC_NMMA": "1"}, {"FFSR_TC_NMMA": "2"}, {"FFSR_TC_NMMA": "3"}, {"FFSR_TC_NMMA": "3", "FFSR_TC_NACC": "3"},
           {"FFSR_TC_NMMA": "3", "FFSR_TC_LEAN0": "1"}, {"FFSR_TC_NMMA": "3", "FFSR_TC_EPI_OWN_FORCE": "1"}]
sizes = [(1, 256, 256), (1, 96, 96)] if len(sys.argv) < 2 else [(1, int(sys.argv[1]), int(sys.argv[2]))]
for (N, H, W) in sizes:
    for ci, co, ks, act, obf in shapes:
        g = torch.Generator().manual_seed(ci + co)
        x = torch.randn(N, H, W, (ci + 7) // 8 * 8, generator=g).to(dev).bfloat16()
        wt = torch.randn(co, ci, ks, ks, generator=g) / (ci * ks * ks) ** 0.5
        eng._w = {"t": _pack_conv(wt).to(dev), "t.b": torch.randn(co, generator=g).to(dev)}
        outs = []
        for cfg in configs:
            for k in ("FFSR_TC_NMMA", "FFSR_TC_NACC", "FFSR_TC_LEAN0", "FFSR_TC_EPI_OWN_FORCE"):
                os.environ.pop(k, None)
            os.environ.update(cfg)
            out = torch.zeros(N, H, W, (co + 7) // 8 * 8, device=dev, dtype=torch.bfloat16 if obf else torch.float32)
            eng.conv(nhwc(x), N, H, W, ci, "t", co, ks, nhwc(out), act=act)
            torch.cuda.synchronize()
            outs.append(out.float())
        ref = outs[0]
        res = []
        for cfg, o in zip(configs[1:], outs[1:]):
            bad = (o != ref)
            res.append(f"{int(bad.sum())}")
        print(f"{H}x{W} {ci:3d}->{co:3d} k{ks} act{act} {'bf16' if obf else 'f32 '}: mismatching values vs NMMA=1 -> " + " | ".join(res), flush=True)
print("configs:", configs[1:])
