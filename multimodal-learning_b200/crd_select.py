"""Selection variant of the CRD distiller -- host-side mirror of the reference's
`MICCAI-2022/CL_utils/CRD_loss.py` (5-arg `CRDLoss(opt, n_data)`, `ContrastLoss_v2`, `weighted_CRDLoss`,
`weighted_ContrastLoss`, single-Linear `Embed`) and `MICCAI-2022/CL_utils/memory_new.py`
(`ContrastMemory_v3`, `ContrastMemory_v2`), the variant `train_test_path_multi_distill.py:202-208,278-284`
actually runs.  Same constructor / forward signatures and buffer names.

    CRDLoss(opt, n_data).forward(epoch, f_s, f_t, idx, contrast_idx=None)            (CRD_loss.py:133,153)
    ContrastMemory_v3(inputSize, outputSize, P, K, T, momentum, select_pos_pairs, P2, select_neg_pairs, K2)
        .forward(epoch, v1, v2, y, idx=None, select_pos_mode="mid") -> (out_v1, out_v2)   (memory_new.py:229,249)
    ContrastLoss_v2(n_data, sample_KD).forward(x, P)                                 (CRD_loss.py:216,221)

How it runs here (all on the GPU, nothing of size [B, K+P, D] is ever written):
  1. `mml_crd_relation_diff`: one pass over the K+P sampled rows of both banks -> diff = t_relation - s_relation
     (memory_new.py:288-292), the only quantity the selection looks at.
  2. selection (:298-361): a sort of the P positive columns and a top-K2 of the K negative columns of `diff`
     (library `torch.sort` / `torch.topk` on a [B, K+P] fp32 matrix); the positions picked in the sorted order come
     from the SAME global numpy RNG calls as the reference (:311,317,321), so seeding numpy gives the same picks.
  3. `mml_crd_fused_loss_grad_multipos` over the P2+K2 selected rows: both ContrastLoss_v2 terms and dL/dv in one
     pass (sample_KD == "False"); or `mml_crd_scores` + `mml_crd_weighted_rows` when the caller wants out_v1/out_v2
     themselves (direct `ContrastMemory_v3.forward`, sample_KD == "True").
  4. the same momentum row update (`mml_crd_memory_update`).
CUDA only; no CPU path.
"""
from __future__ import annotations

import math
import weakref

import numpy as np
import torch
from torch import nn
from torch.autograd.function import once_differentiable

from . import _cabi
from . import crd as _crd
from .crd import AliasMethod, ContrastMemory, Normalize

eps = 1e-7      # CRD_loss.py:5


def sort_columns(diff, col0, n, *, descending, first=None, label0=0):
    """Column numbers of diff[:, col0:col0+n] in sorted order (the first `first` of them), per anchor, + label0:
    `torch.sort(diff[:, col0:col0+n], descending=...)[1][:, :first] + label0` with exact ties ordered by column.
    One CTA per anchor (`mml_crd_sort_columns`: bitonic network in shared memory) for n <= 16384 columns; wider
    problems fall back to the library sort."""
    lib = _cabi.lib()
    B, ld = diff.shape
    first = n if first is None else int(first)
    if n > lib.mml_crd_sort_columns_max():
        part = diff[:, col0:col0 + n]
        order = torch.sort(part, dim=1, descending=descending, stable=True)[1] if first == n else \
            torch.topk(part, first, dim=1, largest=descending, sorted=True)[1]
        return order[:, :first] + label0
    out = torch.empty(B, first, dtype=torch.int64, device=diff.device)
    _cabi.check(lib.mml_crd_sort_columns(_cabi.dptr(diff, torch.float32), B, ld, int(col0), int(n), int(bool(descending)), first,
                                         int(label0), _cabi.dptr(out), first, _cabi.cur_stream(diff.device)),
                "mml_crd_sort_columns")
    return out


def crd_relation_diff(bank1, bank2, v1, v2, idx):
    """diff[b,k] = cos(bank1[idx[b,k]], v1[b]) - cos(bank2[idx[b,k]], v2[b])   (memory_new.py:288-292)."""
    B, D = v1.shape
    diff = torch.empty(idx.shape, dtype=torch.float32, device=v1.device)
    ptr, nbytes = _crd._idx_arg(idx)
    _cabi.check(_cabi.lib().mml_crd_relation_diff(
        _cabi.dptr(bank1), _cabi.dptr(bank2), bank1.shape[0], D, _cabi.dptr(v1), _cabi.dptr(v2), ptr, nbytes, B,
        idx.shape[1], _cabi.dptr(diff), _cabi.cur_stream(v1.device)), "mml_crd_relation_diff")
    return diff


def crd_fused_loss_grad_multipos(bank1, bank2, v1, v2, idx, n_pos, T, Z, n_data, want_out=False):
    """-> (loss[1], g1[B,D], g2[B,D], out_v1 | None, out_v2 | None); the first n_pos columns of idx are positives."""
    B, D = v1.shape
    cols = idx.shape[1]
    dev = v1.device
    ws = _crd._workspace(B, cols, D, dev)
    g1, g2 = torch.empty_like(v1), torch.empty_like(v2)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    out1 = torch.empty(idx.shape, dtype=torch.float32, device=dev) if want_out else None
    out2 = torch.empty(idx.shape, dtype=torch.float32, device=dev) if want_out else None
    ptr, nbytes = _crd._idx_arg(idx)
    if _crd.KERNEL_TIMER is not None:
        _crd.KERNEL_TIMER.start("crd_fused_loss_grad_multipos", dev)
    rc = _cabi.lib().mml_crd_fused_loss_grad_multipos(
        _cabi.dptr(bank1), _cabi.dptr(bank2), bank1.shape[0], D, _cabi.dptr(v1), _cabi.dptr(v2), ptr, nbytes, B, cols,
        int(n_pos), float(T), _cabi.dptr(Z), int(n_data), _cabi.dptr(loss), _cabi.dptr(g1), _cabi.dptr(g2),
        _cabi.dptr(out1), _cabi.dptr(out2), _cabi.dptr(ws), ws.numel(), _cabi.cur_stream(dev))
    if _crd.KERNEL_TIMER is not None:
        _crd.KERNEL_TIMER.stop("crd_fused_loss_grad_multipos", dev)
    _cabi.check(rc, "mml_crd_fused_loss_grad_multipos")
    return loss, g1, g2, out1, out2


class _FusedMultiPosFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v1, v2, mem, sel_idx, n_pos, n_data):
        loss, g1, g2, _, _ = crd_fused_loss_grad_multipos(mem.memory_v1, mem.memory_v2, v1, v2, sel_idx, n_pos, mem._T,
                                                          mem.params[2:4], n_data)
        ctx.save_for_backward(g1, g2)
        return loss.reshape(())                     # CRD_loss.py:241 yields a 0-dim tensor

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loss):
        g1, g2 = ctx.saved_tensors                  # repeatable (train_test_path_multi_distill.py:49-56)
        return grad_loss * g1, grad_loss * g2, None, None, None, None


class ContrastMemory_v3(ContrastMemory):
    """Select the positive and negative pairs simultaneously (memory_new.py:225-397)."""

    _mid_rule = "choice_30_100"                     # memory_new.py:311

    def __init__(self, inputSize, outputSize, P, K, T=0.07, momentum=0.5, select_pos_pairs=True, P2=10,
                 select_neg_pairs=True, K2=512):
        nn.Module.__init__(self)
        self.nLem = outputSize
        self.unigrams = torch.ones(self.nLem)
        self.multinomial = AliasMethod(self.unigrams)
        self.P, self.K, self.P2, self.K2 = P, K, P2, K2
        self.select_pos_pairs = select_pos_pairs
        self.select_neg_pairs = select_neg_pairs
        self.register_buffer('params', torch.tensor([K, T, -1, -1, momentum, P]))
        stdv = 1. / math.sqrt(inputSize / 3)
        self.register_buffer('memory_v1', torch.rand(outputSize, inputSize).mul_(2 * stdv).add_(-stdv))
        self.register_buffer('memory_v2', torch.rand(outputSize, inputSize).mul_(2 * stdv).add_(-stdv))
        self._refresh_scalars()
        self._pending = weakref.WeakSet()

    def _refresh_scalars(self):
        super()._refresh_scalars()
        self._P = int(self.params.detach().cpu()[5].item())

    # ---- selection (memory_new.py:298-361) ----
    def _positive_picks(self, epoch, select_pos_mode):
        """Positions in the sorted order, from the global numpy RNG exactly as the reference draws them."""
        if select_pos_mode == "hard":
            return None
        if select_pos_mode == "mid":
            if self._mid_rule == "choice_30_100":
                return np.random.choice(np.arange(30, 100, 1), self.P2, replace=False)      # memory_new.py:311
            return np.random.randint(50, 100, self.P2)                                      # memory_new.py:152 (v2)
        if select_pos_mode == "random":
            return np.random.randint(0, self.P, self.P2)                                    # :317
        if select_pos_mode == "curriculum":
            interval = 4 - np.ceil(3 * epoch)                                               # :320
            return np.random.randint(50 * (interval - 1), 50 * interval, self.P2)           # :321
        raise NotImplementedError(select_pos_mode)

    def _neg_selected(self):
        if self.select_neg_pairs in ("True", True):
            return True
        if self.select_neg_pairs in ("False", False):
            return False
        raise RuntimeError(f"select_neg_pairs must be 'True' or 'False' (memory_new.py:339,359); got {self.select_neg_pairs!r}")

    def select(self, epoch, v1, v2, idx, select_pos_mode="mid"):
        """-> (sel [B, P2 + (K2|K)] column numbers into idx, sel_idx = idx.gather(1, sel)); column 0 = exact positive."""
        if self.select_pos_pairs is not True:
            raise RuntimeError("select_pos_pairs must be True: the reference's forward needs out_v2_pos (memory_new.py:363)")
        P, K = self._P, self._K
        diff = crd_relation_diff(self.memory_v1, self.memory_v2, v1.detach(), v2.detach(), idx)
        order = sort_columns(diff, 0, P, descending=True)                                   # :303
        picks = self._positive_picks(epoch, select_pos_mode)
        if picks is None:
            sel_pos = order[:, :self.P2].clone()                                            # :307
        else:
            sel_pos = order.index_select(1, torch.as_tensor(np.asarray(picks), dtype=torch.long).to(idx.device))
        sel_pos[:, 0] = 0                                                                   # :325
        if self._neg_selected():
            # ascending order of diff, first K2 (:342-345) == top-K2 smallest, returned sorted
            sel_neg = sort_columns(diff, P, K, descending=False, first=min(self.K2, K), label0=P)
        else:
            sel_neg = torch.arange(P, P + K, device=idx.device).view(1, -1).expand(idx.shape[0], -1)   # :359-361
        sel = torch.cat((sel_pos, sel_neg), 1)
        return sel, idx.gather(1, sel).contiguous()

    def _check_inputs_v3(self, v1, v2, y, idx):
        if not (v1.is_cuda and self.memory_v1.is_cuda):
            raise RuntimeError("ContrastMemory_v3 runs on CUDA tensors only (move the module with .cuda()/.to(device))")
        B = v1.size(0)
        v1, v2, y = _crd._as_f32(v1), _crd._as_f32(v2), _crd._as_i64(y)
        cols = self._K + self._P
        if idx is None:                                                                    # :263-265
            idx = self.multinomial.draw(B * cols, y=y, cols=cols).view(B, -1)
        elif idx.dtype != torch.int32:
            idx = _crd._as_i64(idx)
        idx = idx.contiguous().view(B, cols)       # same RuntimeError as :269 when idx has the wrong width
        return v1, v2, y, idx

    def forward(self, epoch, v1, v2, y, idx=None, select_pos_mode="mid"):
        "v1 is the feature of the student model, v2 refer to the teacher feature."
        v1, v2, y, idx = self._check_inputs_v3(v1, v2, y, idx)
        _, sel_idx = self.select(epoch, v1, v2, idx, select_pos_mode)
        self._ensure_Z(v1, v2, sel_idx)            # mean over the SELECTED scores (:367-374)
        if torch.is_grad_enabled() and (v1.requires_grad or v2.requires_grad):
            undo = _crd._UndoLog()
            self._pending.add(undo)
            out_v1, out_v2 = _crd._ScoresFn.apply(v1, v2, self, sel_idx, undo)
        else:
            o1, o2, _ = _crd.crd_scores(self.memory_v1, self.memory_v2, v1, v2, sel_idx, self._T, Z=self.params[2:4])
            out_v1, out_v2 = o1.unsqueeze(2), o2.unsqueeze(2)
        self._update(v1, v2, y)
        return out_v1, out_v2

    def fused_nce_loss_v2(self, epoch, v1, v2, y, idx, n_data, select_pos_mode):
        """criterion_s(out_v1, P2) + criterion_t(out_v2, P2) of the 5-arg CRDLoss (CRD_loss.py:168-174, sample_KD ==
        "False") without materialising out_v1/out_v2; same side effects as `forward`."""
        v1, v2, y, idx = self._check_inputs_v3(v1, v2, y, idx)
        _, sel_idx = self.select(epoch, v1, v2, idx, select_pos_mode)
        self._ensure_Z(v1, v2, sel_idx)
        loss = _FusedMultiPosFn.apply(v1, v2, self, sel_idx, self.P2, n_data)
        self._update(v1, v2, y)
        return loss


class ContrastMemory_v2(ContrastMemory_v3):
    """memory buffer that supplies large amount of positive and negative samples (memory_new.py:83-222): the positive
    selection of v3, every one of the K negatives kept, and its own "mid" rule (:152)."""

    _mid_rule = "randint_50_100"

    def __init__(self, inputSize, outputSize, P, K, T=0.07, momentum=0.5, select_pos_pairs=True, P2=10):
        super().__init__(inputSize, outputSize, P, K, T, momentum, select_pos_pairs, P2, select_neg_pairs="False", K2=K)


class ContrastMemory_v4(ContrastMemory_v3):
    """`MIA 2022/CL_utils/memory_new.py:398-561`: P2 selected positives + ALL K negatives, the negatives re-weighted by how much
    more similar they look to the student than to the teacher (`relation_diff = s_relation - t_relation + 1`, :493-497).
    Relations here are (bank-1 rows, v1) for the student side and (bank-2 rows, v2) for the teacher side (:462-466), i.e.
    the MIRROR of ContrastMemory_v3's: positives are ordered by (t - s) = -(gap the relation kernel emits), the weight of a
    negative is gap + 1.  Same kernels as v3 (relation gaps, per-anchor sort, scores with autograd, row update); the
    selection / re-weighting glue on [B, P+K] stays host-side PyTorch."""

    _mid_rule = "choice_30_100"                                                             # :478

    def __init__(self, inputSize, outputSize, P, K, T=0.07, momentum=0.5, select_pos_pairs=True, P2=10,
                 select_neg_pairs=False, neg_reweight=True, K2=512):
        super().__init__(inputSize, outputSize, P, K, T, momentum, select_pos_pairs, P2, select_neg_pairs, K2)
        self.neg_reweight = neg_reweight

    def _reweight(self):
        if self.neg_reweight == "True":
            return True
        if self.neg_reweight == "False":
            return False
        raise RuntimeError(f"neg_reweight must be 'True' or 'False' (memory_new.py:493,502); got {self.neg_reweight!r}")

    def _select_v4(self, epoch, v1, v2, idx, select_pos_mode, sign):
        """-> (sel_pos [B, P2] columns, w [B, K] negative weights or None).  `sign` = -1: positives by descending (t - s)
        with this class's relation roles; +1: the v3 / mono roles."""
        if self.select_pos_pairs is not True:
            raise RuntimeError("select_pos_pairs must be True: the reference's forward needs out_v2_pos")
        P = self._P
        diff = crd_relation_diff(self.memory_v1, self.memory_v2, v1.detach(), v2.detach(), idx)
        order = sort_columns(diff, 0, P, descending=(sign > 0))
        picks = self._positive_picks(epoch, select_pos_mode)
        if picks is None:
            sel_pos = order[:, :self.P2].clone()
        else:
            sel_pos = order.index_select(1, torch.as_tensor(np.asarray(picks), dtype=torch.long).to(idx.device))
        sel_pos[:, 0] = 0
        return sel_pos, diff

    def _combine(self, o, sel_pos, w):
        """[B, P+K] scores -> cat(selected positives, (re-weighted) negatives) [B, P2+K, 1]  (:489-507)."""
        P, K = self._P, self._K
        neg = o[:, P:P + K]
        if w is not None:
            neg = neg * w
        return torch.cat((o.gather(1, sel_pos), neg), 1).unsqueeze(2)

    def forward(self, epoch, v1, v2, y, idx=None, select_pos_mode="mid"):
        "v1 is the feature of the student model, v2 refer to the teacher feature."
        v1, v2, y, idx = self._check_inputs_v3(v1, v2, y, idx)
        reweight = self._reweight()
        sel_pos, diff = self._select_v4(epoch, v1, v2, idx, select_pos_mode, sign=-1)
        w = (diff[:, self._P:self._P + self._K] + 1.0) if reweight else None                # s_relation - t_relation + 1
        if not self._z_ready:                                                               # :512-519, over the COMBINED scores
            r1, r2, _ = _crd.crd_scores(self.memory_v1, self.memory_v2, v1.detach(), v2.detach(), idx, self._T)
            self.params[2] = self._combine(r1, sel_pos, w).mean() * self.nLem
            self.params[3] = self._combine(r2, sel_pos, w).mean() * self.nLem
            print("normalization constant Z_v1 is set to {:.1f}".format(self.params[2].item()))
            print("normalization constant Z_v2 is set to {:.1f}".format(self.params[3].item()))
            self._z_ready = True
        if torch.is_grad_enabled() and (v1.requires_grad or v2.requires_grad):
            undo = _crd._UndoLog()
            self._pending.add(undo)
            o1, o2 = _crd._ScoresFn.apply(v1, v2, self, idx, undo)
            o1, o2 = o1.squeeze(2), o2.squeeze(2)
        else:
            o1, o2, _ = _crd.crd_scores(self.memory_v1, self.memory_v2, v1, v2, idx, self._T, Z=self.params[2:4])
        out_v1, out_v2 = self._combine(o1, sel_pos, w), self._combine(o2, sel_pos, w)
        self._update(v1, v2, y)
        return out_v1.contiguous(), out_v2.contiguous()


class ContrastMemory_mono(ContrastMemory_v4):
    """`MIA 2022/CL_utils/memory_new.py:565-698`: one-directional bank -- queries from the student (v2), keys from the
    teacher bank (memory_v1); `forward(epoch, v1 = teacher, v2 = student, y, idx)` returns `(out_v2, memory_v1)`.
    params = [P, K, T, Z_v2, momentum] (:586).  Positives by descending (t - s) with t = (bank-1 rows, v1), s = (bank-2
    rows, v2) -- the relation kernel's own gap; every negative is kept, none re-weighted; 'mid' draws randint(50, 100)."""

    _mid_rule = "randint_50_100"                                                            # :643

    def __init__(self, inputSize, outputSize, P, K, T=0.07, momentum=0.5, select_pos_pairs=True, P2=10,
                 select_neg_pairs=False, neg_reweight=True, K2=512):
        super().__init__(inputSize, outputSize, P, K, T, momentum, select_pos_pairs, P2, select_neg_pairs, neg_reweight, K2)
        del self.params
        self.register_buffer('params', torch.tensor([P, K, T, -1, momentum]))
        self._refresh_scalars()

    def _refresh_scalars(self):
        p = self.params.detach().cpu()
        if p.numel() == 6:                      # called from the parents' constructors before the layout is replaced
            return super()._refresh_scalars()
        self._P, self._K = int(p[0].item()), int(p[1].item())
        self._T = p[2].item()
        self._momentum = p[4].item()
        self._z_ready = bool(p[3].item() > 0)

    def forward(self, epoch, v1, v2, y, idx=None, select_pos_mode="hard"):
        v1, v2, y, idx = self._check_inputs_v3(v1, v2, y, idx)
        sel_pos, _ = self._select_v4(epoch, v1, v2, idx, select_pos_mode, sign=+1)
        if not self._z_ready:                                                               # :668-671
            _, r2, _ = _crd.crd_scores(self.memory_v1, self.memory_v2, v1.detach(), v2.detach(), idx, self._T)
            self.params[3] = self._combine(r2, sel_pos, None).mean() * self.nLem
            print("normalization constant Z_v2 is set to {:.1f}".format(self.params[3].item()))
            self._z_ready = True
        zpair = torch.stack((torch.ones_like(self.params[3]), self.params[3]))             # side 1 is not used
        if torch.is_grad_enabled() and v2.requires_grad:
            undo = _crd._UndoLog()
            self._pending.add(undo)
            proxy = _MonoScores(self, zpair)
            _, o2 = _crd._ScoresFn.apply(v1.detach(), v2, proxy, idx, undo)
            o2 = o2.squeeze(2)
        else:
            _, o2, _ = _crd.crd_scores(self.memory_v1, self.memory_v2, v1, v2, idx, self._T, Z=zpair)
        out_v2 = self._combine(o2, sel_pos, None)
        self._update(v1, v2, y)
        return out_v2.contiguous(), self.memory_v1


class _MonoScores:
    """What `_ScoresFn` reads from a memory module, with the [.., .., Z_v1, Z_v2] parameter layout it expects."""

    def __init__(self, mem, zpair):
        self.memory_v1, self.memory_v2, self._T = mem.memory_v1, mem.memory_v2, mem._T
        self.params = torch.cat((zpair.new_zeros(2), zpair))


class ContrastLoss_v2(nn.Module):
    """supervised contrastive loss (CRD_loss.py:212-252) -- stand-alone form for callers that hold out_v1/out_v2."""

    def __init__(self, n_data, sample_KD):
        super(ContrastLoss_v2, self).__init__()
        self.n_data = n_data
        self.sample_KD = sample_KD

    def forward(self, x, P):
        bsz = x.shape[0]
        N = x.size(1) - P
        m = N
        Pn = 1 / float(self.n_data)
        P_pos = x.narrow(1, 0, P)
        log_D1 = (P_pos / (P_pos + (m * Pn + eps))).log()
        P_neg = x.narrow(1, P, N)
        log_D0 = ((m * Pn) / (P_neg + (m * Pn + eps))).log()
        if self.sample_KD == "False":
            # average of the 1 exact pos. sample and (P-1) relax pos. samples (:241)
            return -((log_D1.squeeze().sum(0) + log_D0.reshape(-1, 1).sum(0)) / bsz).sum(0) / P
        elif self.sample_KD == "True":
            return -(log_D1.squeeze(-1) + log_D0.sum(1)).sum(1) / P                       # :245, per-sample [B]
        raise RuntimeError(f"sample_KD must be 'True' or 'False'; got {self.sample_KD!r}")   # reference: UnboundLocalError


class Embed(_crd.Embed):
    """Embedding module of the variant: ONE Linear + L2 normalisation (CRD_loss.py:256-267)."""

    def __init__(self, dim_in=1024, dim_out=128):
        super().__init__(dim_in, dim_out, layers=1)


class CRDLoss(nn.Module):
    """CRD Loss function, 5-arg selection variant (CRD_loss.py:127-175).

    Args: opt.s_dim / t_dim / feat_dim, opt.nce_p / nce_p2 (candidate / kept positives), opt.nce_k / nce_k2
    (candidate / kept negatives), opt.nce_t, opt.nce_m, opt.select_pos_pairs, opt.select_neg_pairs ("True"/"False"),
    opt.select_pos_mode (hard | mid | random | curriculum), opt.sample_KD ("True"/"False"); n_data = bank rows."""

    def __init__(self, opt, n_data):
        super(CRDLoss, self).__init__()
        self.P = opt.nce_p
        self.P2 = opt.nce_p2
        self.embed_s = Embed(opt.s_dim, opt.feat_dim)
        self.embed_t = Embed(opt.t_dim, opt.feat_dim)
        self.contrast = ContrastMemory_v3(opt.feat_dim, n_data, opt.nce_p, opt.nce_k, opt.nce_t, opt.nce_m,
                                          opt.select_pos_pairs, opt.nce_p2, opt.select_neg_pairs, opt.nce_k2)
        self.criterion_t = ContrastLoss_v2(n_data, sample_KD=opt.sample_KD)
        self.criterion_s = ContrastLoss_v2(n_data, sample_KD=opt.sample_KD)
        self.select_pos_mode = opt.select_pos_mode

    def forward(self, epoch, f_s, f_t, idx, contrast_idx=None):
        """
        f_s / f_t: [batch_size, s_dim / t_dim] student / teacher feature;  idx: [batch_size] dataset indices
        contrast_idx: [batch_size, nce_p + nce_k] (first nce_p columns positives, column 0 the anchor itself) or None
        Returns the contrastive loss: 0-dim (sample_KD "False") or [batch_size] (sample_KD "True").
        """
        f_s, f_t = _crd.embed_pair(self.embed_s, f_s, self.embed_t, f_t)      # the two heads on two streams
        if self.criterion_s.sample_KD == "False" and self.criterion_s.n_data == self.criterion_t.n_data:
            return self.contrast.fused_nce_loss_v2(epoch, f_s, f_t, idx, contrast_idx, self.criterion_s.n_data,
                                                   self.select_pos_mode)
        out_s, out_t = self.contrast(epoch, f_s, f_t, idx, contrast_idx, self.select_pos_mode)
        return self.criterion_s(out_s, self.P2) + self.criterion_t(out_t, self.P2)


class weighted_ContrastLoss(nn.Module):
    """contrastive loss with per-sample weights (CRD_loss.py:53-81)."""

    def __init__(self, n_data):
        super(weighted_ContrastLoss, self).__init__()
        self.n_data = n_data

    def forward(self, x, sample_weights):
        bsz = x.shape[0]
        m = x.size(1) - 1
        Pn = 1 / float(self.n_data)
        P_pos = x.select(1, 0)
        log_D1 = (P_pos / (P_pos + (m * Pn + eps))).log()
        P_neg = x.narrow(1, 1, m)
        log_D0 = ((m * Pn) / (P_neg + (m * Pn + eps))).log()
        return -torch.sum(sample_weights * (log_D1 + log_D0.view(bsz, -1).sum(1, keepdims=True))) / bsz


class weighted_CRDLoss(nn.Module):
    """CRD loss whose two sides are gated per sample by which of two task losses is larger (CRD_loss.py:8-50)."""

    def __init__(self, opt, n_data):
        super(weighted_CRDLoss, self).__init__()
        self.embed_s = Embed(opt.s_dim, opt.feat_dim)
        self.embed_t = Embed(opt.t_dim, opt.feat_dim)
        self.contrast = ContrastMemory(opt.feat_dim, n_data, opt.nce_k, opt.nce_t, opt.nce_m)
        self.criterion_t = weighted_ContrastLoss(n_data)
        self.criterion_s = weighted_ContrastLoss(n_data)

    def forward(self, f_s, f_t, loss_s, loss_t, idx, contrast_idx=None):
        f_s, f_t = _crd.embed_pair(self.embed_s, f_s, self.embed_t, f_t)      # the two heads on two streams
        out_s, out_t = self.contrast(f_s, f_t, idx, contrast_idx)
        s_weight = torch.where(loss_s > loss_t, 1.0, 0.0)
        t_weight = torch.where(loss_t > loss_s, 1.0, 0.0)
        return self.criterion_s(out_s, s_weight) + self.criterion_t(out_t, t_weight)
