"""Drop-in for the reference's `MIA 2022/CL_utils/CRD_loss_v2.py`: `CRDLoss(opt, n_data)` over ContrastMemory_v4 and
`CRDLoss_v2(opt, n_data)` over ContrastMemory_mono, `forward(epoch, f_s, f_t, idx, contrast_idx)`."""
from multimodal_learning_b200.crd import Normalize  # noqa: F401
from multimodal_learning_b200.crd_loss_v2 import ContrastLoss_v2, CRDLoss, CRDLoss_v2, Embed, eps  # noqa: F401
