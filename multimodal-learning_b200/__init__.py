"""multimodal-learning_b200: B200-native (sm_100a) fusion + CRD distillation hot path.

Drop-in mirrors of the reference's modules (CityU-AIM-Group/MultiModal-learning,
`MICCAI-2022/fusion.py`, `MICCAI-2022/CL_utils/CRD_criterion.py`, `CL_utils/CRD_loss.py`, `CL_utils/memory_new.py`,
`MICCAI-2022/KD_loss.py`)
over hand-written CUDA kernels reached through the C ABI in `include/mml_b200.h`.
The directory name has a hyphen (task-mandated); import it as `multimodal_learning_b200`
(the repo-root shim `multimodal_learning_b200.py` registers it), or put
`multimodal-learning_b200/dropin` on PYTHONPATH to shadow the reference's own
`fusion`, `KD_loss` and `CL_utils.CRD_criterion` modules unchanged.
"""
from . import _cabi
from ._cabi import check_device_errors, device_error_flags
from .crd import (AliasMethod, ContrastLoss, ContrastMemory, CRDLoss, Embed, Normalize)
from . import crd_knn, crd_loss_v2, crd_select
from .crd_select import ContrastLoss_v2, ContrastMemory_mono, ContrastMemory_v2, ContrastMemory_v3, ContrastMemory_v4
from .fusion import BilinearFusion, PolynomialFusion, TrilinearFusion_A, TrilinearFusion_B, init_max_weights, kron_linear
from .graphed import GraphedTrainStep
from .kd_loss import DistillKL
from .sampler import InstanceSampler

__all__ = ["BilinearFusion", "PolynomialFusion", "TrilinearFusion_A", "TrilinearFusion_B", "init_max_weights", "kron_linear", "AliasMethod", "ContrastLoss", "ContrastMemory", "CRDLoss", "Embed", "Normalize", "DistillKL", "GraphedTrainStep", "InstanceSampler", "crd_select", "crd_loss_v2", "crd_knn", "check_device_errors", "device_error_flags", "ContrastLoss_v2", "ContrastMemory_v2", "ContrastMemory_v3", "ContrastMemory_v4", "ContrastMemory_mono", "_cabi"]
