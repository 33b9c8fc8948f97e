"""Drop-in for the reference's `CL_utils/CRD_loss.py` (train_test_path_multi_distill.py:22, :202):
`CRDLoss(opt, n_data).forward(epoch, f_s, f_t, idx, contrast_idx)`."""
from multimodal_learning_b200.crd import ContrastLoss, Normalize  # noqa: F401
from multimodal_learning_b200.crd_select import (ContrastLoss_v2, CRDLoss, Embed, eps, weighted_ContrastLoss,  # noqa: F401
                                                  weighted_CRDLoss)
