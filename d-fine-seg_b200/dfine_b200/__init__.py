"""dfine_b200 -- B200-native (sm_100a) decoder hot path of D-FINE-seg.

Python here is host plumbing around the C-ABI library libdfine_b200.so
(include/dfine_b200.h): multi-scale deformable attention forward / backward, the FDR box
decode, the tcgen05 mask-prototype GEMM (forward and backward) and the criterion's Hungarian
assignment on the device.  No CPU fallback exists.
"""
from . import _lib, build, criterion, grad_sync, ops
from ._lib import DfineB200Error, library_path
from .criterion import patch_criterion, unpatch_criterion
from .modules import GraphedInference, Integral, LazyMaskLogits, MSDeformableAttention, patch_model, unpatch_model
from .ops import (fdr_decode, fdr_integral, fdr_project, mask_logits, msda_core, msda_fused,
                  msda_fused_packed)

__all__ = [
    "MSDeformableAttention", "Integral", "patch_model", "unpatch_model", "GraphedInference", "LazyMaskLogits", "msda_core",
    "msda_fused", "msda_fused_packed", "fdr_project", "fdr_integral", "fdr_decode", "mask_logits", "library_path",
    "DfineB200Error", "ops", "build", "grad_sync", "criterion", "patch_criterion", "unpatch_criterion",
]
__version__ = "0.1.0"
