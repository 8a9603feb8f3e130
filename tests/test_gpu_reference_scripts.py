"""The drop-in boundary exercised by the REFERENCE'S OWN code on a GPU: unmodified scripts of the reference (the git-ignored
baseline/_ref copy that tools/install_reference.py makes, or /root/reference in the build container) run through
``python -m isr_b200.install``, which pre-seeds ``src.models.enhanced_fusion_v2`` with the sm_100a module."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = next((p for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")
            if os.path.isfile(os.path.join(p, "scripts", "test_cached_training.py"))), None)
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(REF is None, reason="no copy of the reference on this machine (baseline/_ref)")]


def _run(argv, timeout=900):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    return subprocess.run([sys.executable, "-m", "isr_b200.install"] + argv, cwd=REF, env=env, capture_output=True, text=True, timeout=timeout)


def test_reference_cached_training_script_passes_on_the_installed_module():
    """scripts/test_cached_training.py of the reference, unmodified: its CachedSRDataset over mock pickles, the model in cached
    mode, forward_with_precomputed, gradient flow to all parameters, a torch.optim.Adam step, forward() raising in cached mode."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = _run([os.path.join(REF, "scripts", "test_cached_training.py")])
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert out.count("[PASSED]") >= 7 and "[FAILED]" not in out, out[-3000:]
    assert "Params with gradients: 198/198" in out, out[-3000:]


def test_reference_validation_loop_over_its_own_dataset(tmp_path):
    """A test.py / validate_epoch-style loop written against the reference's names only (src.data.CachedSRDataset, DataLoader,
    src.models.CompleteEnhancedFusionSR, src.utils.metrics): train.py:415-515 in cached mode on mock pickles."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    script = tmp_path / "val_loop.py"
    script.write_text(textwrap.dedent(f"""
        import sys, tempfile
        from pathlib import Path
        sys.path.insert(0, {REF!r})
        import torch
        from torch.utils.data import DataLoader
        from src.data import CachedSRDataset
        from src.models import CompleteEnhancedFusionSR
        from src.utils.metrics import calculate_psnr, calculate_ssim
        import isr_b200
        assert CompleteEnhancedFusionSR is isr_b200.CompleteEnhancedFusionSR, "install did not take effect"
        d = Path(tempfile.mkdtemp())
        g = torch.Generator().manual_seed(0)
        for i in range(3):
            stem = f"img_{{i:03d}}"
            hr = torch.rand(3, 128, 128, generator=g)
            torch.save({{"outputs": {{"drct": (hr + 0.02 * torch.randn(3, 128, 128, generator=g)).clamp(0, 1).unsqueeze(0)}},
                        "features": {{"drct": torch.randn(1, 180, 32, 32, generator=g)}},
                        "lr": torch.nn.functional.avg_pool2d(hr, 4), "hr": hr}}, d / f"{{stem}}_drct_part.pt")
            torch.save({{"outputs": {{k: (hr + 0.02 * torch.randn(3, 128, 128, generator=g)).clamp(0, 1).unsqueeze(0) for k in ("grl", "nafnet")}},
                        "features": {{"grl": torch.randn(1, 180, 32, 32, generator=g), "nafnet": torch.randn(1, 64, 32, 32, generator=g)}}}},
                       d / f"{{stem}}_rest_part.pt")
            torch.save({{"outputs": {{"mamba": (hr + 0.02 * torch.randn(3, 128, 128, generator=g)).clamp(0, 1).unsqueeze(0).half()}},
                        "features": {{"mamba": torch.randn(1, 180, 32, 32, generator=g).half()}}}}, d / f"{{stem}}_mamba_part.pt")
        ds = CachedSRDataset(str(d), augment=False, repeat_factor=1)
        loader = DataLoader(ds, batch_size=1, shuffle=False, num_workers=0)
        model = CompleteEnhancedFusionSR(expert_ensemble=None).cuda().eval()
        psnrs = []
        with torch.no_grad():
            for batch in loader:
                lr = batch["lr"].cuda()
                imgs = {{k: v.cuda() for k, v in batch["expert_imgs"].items()}}
                feats = {{k: v.cuda() for k, v in batch["expert_feats"].items()}}
                sr = model.forward_with_precomputed(lr, imgs, feats).clamp(0, 1)
                psnrs.append(float(calculate_psnr(sr, batch["hr"].cuda(), crop_border=4)))
                float(calculate_ssim(sr, batch["hr"].cuda(), crop_border=4))
        assert len(psnrs) == 3 and all(10.0 < p < 60.0 for p in psnrs), psnrs     # random-init weights: a sanity band only
        print("VALIDATION LOOP OK", psnrs)
        """))
    r = _run([str(script)])
    assert r.returncode == 0 and "VALIDATION LOOP OK" in r.stdout, (r.stdout + r.stderr)[-3000:]
