"""Install the sm_100a fusion module into an UNMODIFIED checkout of the reference.

The reference scripts put their own root first on ``sys.path`` (train.py:43-44), so a
PYTHONPATH overlay cannot shadow ``src.models.enhanced_fusion_v2``.  Instead the module object
is pre-seeded in ``sys.modules`` before the script runs:

    python -m isr_b200.install /path/to/reference/test.py --test_dir ... --save_dir ...

or, from Python:  ``import isr_b200.install as I; I.install()`` before importing ``src.models``.

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


def _legacy(name):
    def __init__(self, *a, **k):
        raise NotImplementedError(f"{name} is legacy code that CompleteEnhancedFusionSR never instantiates; "
                                  "it is not part of the sm_100a hot path")
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


def install() -> types.ModuleType:
    mod = sys.modules.get(TARGET)
    if mod is None or getattr(mod, "CompleteEnhancedFusionSR", None) is not CompleteEnhancedFusionSR:
        mod = make_module()
        sys.modules[TARGET] = mod
    return mod


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m isr_b200.install <reference script.py> [script args...]")
    install()
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
