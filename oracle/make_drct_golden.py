"""Generate tests/golden/drct_small.npz from the REFERENCE's own DRCT class (run in the build container):

    PYTHONPATH=/root/reference python oracle/make_drct_golden.py

Reduced configuration (2 RDGs, window 8, img_size 16) with every channel count / head rule of DRCT-L; weights come from
``oracle.drct_oracle.synth_state_dict`` so no checkpoint is needed.  ``timm`` is not installed here: the two helpers the
reference takes from it (``to_2tuple``, ``trunc_normal_``; src/models/drct/__init__.py:20) are shimmed, the class itself
runs unmodified.
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import drct_oracle as DO  # noqa: E402

SMALL = dict(embed_dim=180, n_rdg=2, window=8, num_heads=6, gc=32, mlp_ratio=2, num_feat=64, img_size=16)


def reference_class():
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models, models.layers = models, layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    with contextlib.redirect_stdout(io.StringIO()):
        from src.models.drct import DRCT
    return DRCT


def build_reference(cfg=SMALL, seed=3):
    DRCT = reference_class()
    m = DRCT(upscale=4, in_chans=3, img_size=cfg["img_size"], window_size=cfg["window"], compress_ratio=3, squeeze_factor=30,
             conv_scale=0.01, overlap_ratio=0.5, img_range=1.0, depths=[6] * cfg["n_rdg"], embed_dim=cfg["embed_dim"],
             num_heads=[cfg["num_heads"]] * cfg["n_rdg"], mlp_ratio=cfg["mlp_ratio"], upsampler="pixelshuffle",
             resi_connection="1conv").eval()
    shapes = DO.state_shapes(**cfg)
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert list(ref_shapes.items()) == [(k, s) for k, (s, _) in shapes.items()], "state_shapes() differs from the reference"
    sd = DO.synth_state_dict(shapes, seed=seed, img_size=cfg["img_size"])
    for k, v in m.state_dict().items():
        if shapes[k][1] != "float":
            assert torch.equal(v, sd[k].to(v.dtype)), k                 # our index / mask restatements equal the buffers
    m.load_state_dict(sd, strict=True)
    return m, sd


def main():
    m, sd = build_reference()
    g = torch.Generator().manual_seed(77)
    x = torch.rand(1, 3, 16, 24, generator=g)
    feats = {}
    hook = m.conv_after_body.register_forward_hook(lambda mod, i, o: feats.__setitem__("f", o.detach()))
    with torch.no_grad():
        y = m(x)
    hook.remove()
    out = os.path.join(ROOT, "tests", "golden", "drct_small.npz")
    np.savez_compressed(out, x=x.numpy(), y=y.numpy(), feat_sub=feats["f"].numpy().reshape(-1)[::5],
                        cfg=json.dumps(SMALL), seed=3,
                        full_shapes_sha=np.frombuffer(DO_shapes_digest().encode(), dtype=np.uint8))
    print("wrote", out, y.shape, float(y.mean()), os.path.getsize(out))


def DO_shapes_digest():
    """Digest of the names / shapes of the FULL DRCT-L state_dict as the reference builds it (create_drct_model)."""
    import hashlib
    with contextlib.redirect_stdout(io.StringIO()):
        from src.models.drct import create_drct_model
        full = create_drct_model()
    ref = [(k, tuple(v.shape)) for k, v in full.state_dict().items()]
    ours = [(k, s) for k, (s, _) in DO.state_shapes().items()]
    assert ref == ours, "state_shapes() of DRCT-L differs from create_drct_model()"
    return hashlib.sha256(repr(ref).encode()).hexdigest()


if __name__ == "__main__":
    main()
