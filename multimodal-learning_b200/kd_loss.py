"""DistillKL -- host-side mirror of the reference's logit distillation loss (`MICCAI-2022/KD_loss.py:7-17`).

KL(softmax(y_t / T) || softmax(y_s / T)) * T^2 / batch on [B, C] logits with C = 3 classes: a few hundred bytes of work
per step, so it stays plain PyTorch on whatever device the logits live on; the drop-in keeps the
reference's constructor and call signature."""
import torch
from torch import nn


class DistillKL(nn.Module):
    """Hinton-style knowledge distillation on temperature-softened logits."""

    def __init__(self, T):
        super().__init__()
        self.T = T

    def forward(self, y_s, y_t):
        batch = y_s.shape[0]
        student_log_prob = torch.log_softmax(torch.div(y_s, self.T), dim=1)
        teacher_prob = torch.softmax(torch.div(y_t, self.T), dim=1)
        # the reference's `size_average=False` is today's reduction="sum" (KD_loss.py:16)
        divergence = nn.functional.kl_div(student_log_prob, teacher_prob, reduction="sum")
        return divergence * self.T ** 2 / batch
