"""DistillKL -- mirror of the reference's `MICCAI-2022/KD_loss.py:7-17`.

Logit KL on [B, C] with C = 3 classes: negligible work, stays host-side PyTorch
(SURVEY.md §8a row a14)."""
import torch.nn as nn
import torch.nn.functional as F


class DistillKL(nn.Module):
    """Distilling the Knowledge in a Neural Network"""

    def __init__(self, T):
        super(DistillKL, self).__init__()
        self.T = T

    def forward(self, y_s, y_t):
        log_p_s = F.log_softmax(y_s / self.T, dim=1)
        p_t = F.softmax(y_t / self.T, dim=1)
        # size_average=False of the reference == reduction='sum'
        return F.kl_div(log_p_s, p_t, reduction='sum') * (self.T ** 2) / y_s.shape[0]
