"""The sys.modules pre-seeding that drops the module into an unmodified reference checkout."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_install_exports_reference_names():
    import isr_b200
    from isr_b200 import install as I
    mod = I.install()
    assert sys.modules["src.models.enhanced_fusion_v2"] is mod
    for n in ("CompleteEnhancedFusionSR", "create_enhanced_fusion", "AdaptiveFrequencyDecomposition",
              "CrossBandAttention", "CollaborativeFeatureLearning", "MultiResolutionFusion", "DynamicExpertSelector"):
        assert hasattr(mod, n), n
    assert mod.CompleteEnhancedFusionSR is isr_b200.CompleteEnhancedFusionSR
    with pytest.raises(NotImplementedError):
        mod.CrossBandAttention()
    del sys.modules["src.models.enhanced_fusion_v2"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout only exists in the build container")
def test_reference_package_import_resolves_to_our_module():
    code = (
        "import sys; sys.path.insert(0, %r); import isr_b200, isr_b200.install as I; I.install();"
        "sys.path.insert(0, %r);"
        "from src.models import CompleteEnhancedFusionSR as A, create_enhanced_fusion;"
        "from src.models.enhanced_fusion_v2 import CompleteEnhancedFusionSR as B2;"
        "assert A is isr_b200.CompleteEnhancedFusionSR and B2 is A;"
        "m = create_enhanced_fusion(None); assert m.cached_mode and len(m.state_dict()) == 226; print('ok')"
    ) % (ROOT, REF)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_install_fused_losses_module():
    from isr_b200 import install as I
    from isr_b200 import losses as FL
    I.install(losses=True)
    mod = sys.modules["src.losses"]
    assert mod.CombinedLoss is FL.CombinedLoss and mod.PYWT_AVAILABLE is True
    crit = mod.CombinedLoss()
    crit.set_weights({"l1": 0.6, "swt": 0.25, "fft": 0.1, "ssim": 0.05})
    assert crit.current_stage == 3 and crit.weights["swt"] == 0.25
    with pytest.raises(NotImplementedError):
        mod.VGGPerceptualLoss()
    del sys.modules["src.losses"], sys.modules["src.models.enhanced_fusion_v2"]
