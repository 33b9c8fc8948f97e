"""Drop-in for the reference's `CL_utils/CRD_criterion.py` (train_test_MT.py:22)."""
from multimodal_learning_b200.crd import (AliasMethod, ContrastLoss, ContrastMemory, CRDLoss, Embed,  # noqa: F401
                                           Normalize, eps)
