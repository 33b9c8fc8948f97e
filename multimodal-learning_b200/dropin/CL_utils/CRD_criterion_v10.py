"""Drop-in for the reference's `MIA 2023/stage2_unimodal_student/CL_utils/CRD_criterion_v10.py`
(`train_test_path_multi_distill.py:24`): `CRDLoss(opt, n_data, train_class_idx).forward(sample_weights, f_s, f_t, batch_label,
idx, contrast_idx)` with KNN ("neighbors") or class-centre ("centers") positives."""
from multimodal_learning_b200.crd_knn import (ContrastLoss, ContrastLoss_v2, ContrastMemory, CRDLoss, Embed, Normalize,  # noqa: F401
                                               eps)
