"""Drop-in for the reference's `KD_loss.py` (train_test_path_multi_distill.py:23)."""
from multimodal_learning_b200.kd_loss import DistillKL  # noqa: F401
