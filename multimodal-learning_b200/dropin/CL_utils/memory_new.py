"""Drop-in for the reference's `CL_utils/memory_new.py` (imported by CL_utils/CRD_loss.py:3 and, in the MIA 2022 tree,
by CL_utils/CRD_loss_v2.py:8)."""
from multimodal_learning_b200.crd import AliasMethod, ContrastMemory  # noqa: F401
from multimodal_learning_b200.crd_select import (ContrastMemory_mono, ContrastMemory_v2, ContrastMemory_v3,  # noqa: F401
                                                  ContrastMemory_v4)  # noqa: F401
