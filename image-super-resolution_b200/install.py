"""Install the sm_100a fusion module into an UNMODIFIED checkout of the reference.

The reference scripts put their own root first on ``sys.path`` (train.py:43-44), so a
PYTHONPATH overlay cannot shadow ``src.models.enhanced_fusion_v2``.  Instead the module object
is pre-seeded in ``sys.modules`` before the script runs:

    python -m isr_b200.install /path/to/reference/test.py --test_dir ... --save_dir ...

or, from Python:  ``import isr_b200.install as I; I.install()`` before importing ``src.models``.

``--fused-losses`` (``install(losses=True)``) additionally pre-seeds ``src.losses`` with the fused
L1 / SWT / FFT / SSIM ``CombinedLoss`` (``from src.losses import CombinedLoss, PYWT_AVAILABLE``,
train.py:549), which also removes the reference's eager VGG19 download (perceptual_loss.py:1122-1125)
that makes ``train.py`` unstartable offline.

``--shard-loader`` (``install(loader=True)``) pre-seeds ``src.data.cached_dataset`` so that
``from src.data import CachedSRDataset, create_cached_dataloader`` (train.py:592, scripts/validate.py:287) resolve to
``isr_b200.cache``: the cache directory the caller names is packed once into a flat shard and batches arrive as device
tensors (one H2D copy + one kernel per batch) instead of three unpickles per sample in DataLoader workers.

Exports the names the reference re-exports from that module (src/models/__init__.py:50-58).
The four legacy classes are dead code upstream (never instantiated by CompleteEnhancedFusionSR,
SURVEY §2.1 #7); they are importable here and raise on construction.
"""
import runpy
import sys
import types

from .fusion import CompleteEnhancedFusionSR, create_enhanced_fusion
from .modules import DynamicExpertSelector

TARGET = "src.models.enhanced_fusion_v2"
LOSSES_TARGET = "src.losses"
LOADER_TARGET = "src.data.cached_dataset"


def _legacy(name):
    def __init__(self, *a, **k):
        raise NotImplementedError(f"{name} is not used by the cached-fusion hot path (legacy fusion classes / "
                                  "loss components outside the stage-1/2/3 curriculum); it is not built for sm_100a")
    return type(name, (), {"__init__": __init__})


def make_module() -> types.ModuleType:
    mod = types.ModuleType(TARGET)
    mod.__doc__ = "sm_100a drop-in for the reference's enhanced_fusion_v2 (isr_b200)"
    mod.CompleteEnhancedFusionSR = CompleteEnhancedFusionSR
    mod.create_enhanced_fusion = create_enhanced_fusion
    mod.DynamicExpertSelector = DynamicExpertSelector
    for n in ("AdaptiveFrequencyDecomposition", "CrossBandAttention", "CollaborativeFeatureLearning",
              "MultiResolutionFusion"):
        setattr(mod, n, _legacy(n))
    return mod


def make_losses_module() -> types.ModuleType:
    from . import losses as FL
    mod = types.ModuleType(LOSSES_TARGET)
    mod.__doc__ = "sm_100a fused stage-1/2/3 losses (isr_b200.losses) behind the reference's src.losses names"
    for n in ("CombinedLoss", "L1Loss", "SSIMLoss", "FFTLoss", "SWTLoss", "PYWT_AVAILABLE", "LPIPS_AVAILABLE",
              "CLIP_AVAILABLE"):
        setattr(mod, n, getattr(FL, n))
    for n in ("L2Loss", "CharbonnierLoss", "VGGPerceptualLoss", "VGGFeatureExtractor", "EdgeLoss", "CLIPPerceptualLoss"):
        setattr(mod, n, _legacy(n))        # outside the stage-1/2/3 curriculum: importable, raise on construction
    return mod


def make_loader_module() -> types.ModuleType:
    import torch
    from . import cache as CA
    mod = types.ModuleType(LOADER_TARGET)
    mod.__doc__ = "flat-shard loader (isr_b200.cache) behind the reference's src.data.cached_dataset names"
    mod.CachedSRDataset = CA.ShardDataset

    def create_cached_dataloader(feature_dir, batch_size=16, num_workers=4, augment=True, repeat_factor=20, pin_memory=True,
                                 persistent_workers=True, prefetch_factor=4, load_features=True):
        """Reference signature (cached_dataset.py:285-296).  On a CUDA machine the batches are produced on the current
        device; the DataLoader-only arguments are accepted and unused."""
        if not torch.cuda.is_available():
            raise RuntimeError("the shard loader (sm_100a build) produces device batches: there is no CPU path "
                               "(use isr_b200.cache.ShardDataset for host-side samples)")
        dev = torch.device("cuda", torch.cuda.current_device())
        return CA.create_cached_dataloader(feature_dir, batch_size=batch_size, augment=augment, repeat_factor=repeat_factor,
                                           load_features=load_features, device=dev)
    mod.create_cached_dataloader = create_cached_dataloader
    return mod


def install(losses: bool = False, loader: bool = False) -> types.ModuleType:
    mod = sys.modules.get(TARGET)
    if mod is None or getattr(mod, "CompleteEnhancedFusionSR", None) is not CompleteEnhancedFusionSR:
        mod = make_module()
        sys.modules[TARGET] = mod
    if losses and LOSSES_TARGET not in sys.modules:
        sys.modules[LOSSES_TARGET] = make_losses_module()
    if loader and LOADER_TARGET not in sys.modules:
        sys.modules[LOADER_TARGET] = make_loader_module()
    return mod


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    fused = shard = False
    while argv and argv[0] in ("--fused-losses", "--shard-loader"):
        fused, shard = fused or argv[0] == "--fused-losses", shard or argv[0] == "--shard-loader"
        argv = argv[1:]
    if not argv:
        raise SystemExit("usage: python -m isr_b200.install [--fused-losses] [--shard-loader] <reference script.py> [script args...]")
    install(losses=fused, loader=shard)
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
