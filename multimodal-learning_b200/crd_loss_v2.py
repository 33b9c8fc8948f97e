"""Mirror of the reference's `MIA 2022/CL_utils/CRD_loss_v2.py`: the re-weighted-negative CRD loss (`CRDLoss`, :13-55,
over ContrastMemory_v4) and the one-directional one (`CRDLoss_v2`, :58-104, over ContrastMemory_mono).  Same class names,
constructor arguments, forward signatures and side effects; the bank work runs in the CUDA kernels behind
`crd_select.ContrastMemory_v4 / ContrastMemory_mono`."""
from __future__ import annotations

from torch import nn

from . import crd as _crd
from .crd import Normalize
from .crd_select import ContrastLoss_v2, ContrastMemory_mono, ContrastMemory_v3, ContrastMemory_v4, Embed, eps  # noqa: F401


class CRDLoss(nn.Module):
    """CRD Loss function with P2 selected positives and re-weighted negatives (CRD_loss_v2.py:13-55).

    Args: as `crd_select.CRDLoss`, plus opt.neg_reweight ("True": weight every negative by s_relation - t_relation + 1;
    "False": keep the scores as they are)."""

    def __init__(self, opt, n_data):
        super(CRDLoss, self).__init__()
        self.P = opt.nce_p
        self.P2 = opt.nce_p2
        self.embed_s = Embed(opt.s_dim, opt.feat_dim)
        self.embed_t = Embed(opt.t_dim, opt.feat_dim)
        self.contrast = ContrastMemory_v4(opt.feat_dim, n_data, opt.nce_p, opt.nce_k, opt.nce_t, opt.nce_m,
                                          opt.select_pos_pairs, opt.nce_p2, opt.select_neg_pairs, opt.neg_reweight,
                                          opt.nce_k2)
        self.criterion_t = ContrastLoss_v2(n_data, sample_KD=opt.sample_KD)
        self.criterion_s = ContrastLoss_v2(n_data, sample_KD=opt.sample_KD)
        self.select_pos_mode = opt.select_pos_mode

    def forward(self, epoch, f_s, f_t, idx, contrast_idx=None):
        """f_s / f_t: [batch_size, s_dim / t_dim]; idx: [batch_size]; contrast_idx: [batch_size, nce_p + nce_k] or None."""
        f_s, f_t = _crd.embed_pair(self.embed_s, f_s, self.embed_t, f_t)      # the two heads on two streams
        out_s, out_t = self.contrast(epoch, f_s, f_t, idx, contrast_idx, self.select_pos_mode)
        s_loss = self.criterion_s(out_s, self.P2)
        t_loss = self.criterion_t(out_t, self.P2)
        return s_loss + t_loss


class CRDLoss_v2(nn.Module):
    """One-directional contrastive KD (CRD_loss_v2.py:58-104): the teacher feature already has feat_dim columns and is only
    L2-normalised; the student is embedded and used as the query against the teacher bank."""

    def __init__(self, opt, n_data):
        super(CRDLoss_v2, self).__init__()
        self.P = opt.nce_p
        self.P2 = opt.nce_p2
        self.select_pos_mode = opt.select_pos_mode
        self.l2norm = Normalize(2)
        self.embed_s = Embed(opt.s_dim, opt.feat_dim)
        self.contrast = ContrastMemory_mono(opt.feat_dim, n_data, opt.nce_p, opt.nce_k, opt.nce_t, opt.nce_m,
                                            opt.select_pos_pairs, opt.nce_p2, opt.select_neg_pairs, opt.neg_reweight,
                                            opt.nce_k2)
        self.criterion_s = ContrastLoss_v2(n_data, sample_KD=opt.sample_KD)

    def forward(self, epoch, f_s, f_t, idx, contrast_idx=None):
        f_t = f_t.clone().detach()
        f_s = self.embed_s(f_s)
        f_t = self.l2norm(f_t)
        out_s, self.memory_t = self.contrast(epoch, f_t, f_s, idx, contrast_idx, self.select_pos_mode)   # :100
        return self.criterion_s(out_s, self.P2)
