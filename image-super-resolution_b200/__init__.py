"""B200-native FreqFusionSR fusion hot path (sm_100a kernels behind the reference's nn.Module API).

Import as ``isr_b200`` (see ``isr_b200.py`` at the repo root: the directory name carries a
hyphen, so it is registered under that alias).
"""
from .fusion import CompleteEnhancedFusionSR, create_enhanced_fusion, EXPERT_ORDER  # noqa: F401
from .modules import DynamicExpertSelector  # noqa: F401
from . import torch_ops  # noqa: F401  (registers torch.ops.ffsr.*: the torch.library face of the C ABI)

__all__ = ["CompleteEnhancedFusionSR", "create_enhanced_fusion", "DynamicExpertSelector", "EXPERT_ORDER"]
