"""Mirror of the reference's `MIA 2023/stage2_unimodal_student/CL_utils/CRD_criterion_v10.py` (what that tree's
`train_test_path_multi_distill.py:24,258-262` runs): CRD whose positives are the query's nearest same-class neighbours in
the FULL memory bank ("neighbors", :69-80 / :108-116) or class centres ("centers", :81-106).

The reference computes `sklearn.cosine_similarity(anchor rows, whole bank)` on the CPU every step, multiplies by a class
mask and sorts all n columns.  Here the neighbour search is `mml_crd_knn_positives` (one TF32 tcgen05 pass over the bank
with a fused per-anchor candidate filter, exact fp32 re-score, csrc/crd_knn.cu); the scores, their autograd and the bank
update are the kernels of `crd.ContrastMemory`.  Same class names, constructor arguments, forward signatures, state_dict
and side effects."""
from __future__ import annotations

import math
import weakref

import numpy as np
import torch
from torch import nn

from . import _cabi
from . import crd as _crd
from . import crd_kmeans as _kmeans
from .crd import Normalize
from .crd_select import Embed  # single Linear + L2 (CRD_criterion_v10.py:316-329)  # noqa: F401

eps = 1e-7


def knn_inv_norms(bank, out=None, rows=None):
    """1 / |row| of every bank row (rows None) or of the listed rows, written into `out` [n] fp32 (allocated when None)."""
    n, D = bank.shape
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=bank.device)
    _cabi.check(_cabi.lib().mml_crd_knn_inv_norms(
        _cabi.dptr(bank, torch.float32), n, D, _cabi.dptr(rows, torch.int64) if rows is not None else None,
        0 if rows is None else rows.numel(), _cabi.dptr(out, torch.float32), _cabi.cur_stream(bank.device)), "mml_crd_knn_inv_norms")
    return out


def knn_positives(bank, row_labels, anchor_rows, anchor_labels, num_pos, *, n_classes=0, inv_norms=None, queries=None,
                  exact_only=False, return_flags=False):
    """-> (neighbors int64 [B, P], similarity fp32 [B, P]) of CRD_criterion_v10.py:71-76 for one bank.
    row_labels int32 [n] (class of every bank row), anchor_rows int64 [B] (the query's own row), anchor_labels int64 [B].
    n_classes in 1..3 promises that every label lies in [0, n_classes) (faster class mask); 0 = arbitrary labels.
    inv_norms: optional fp32 [n] = 1 / |row| kept up to date by the caller (`knn_inv_norms`); None = computed here.
    queries: optional fp32 [B, D] explicit query vectors (then `anchor_rows` may be None): the sharded bank's case."""
    if not bank.is_cuda:
        raise RuntimeError("knn_positives runs on CUDA tensors only")
    n, D = bank.shape
    B = anchor_rows.numel() if queries is None else queries.shape[0]
    dev = bank.device
    lib = _cabi.lib()
    if num_pos > lib.mml_crd_knn_max_positives():
        raise NotImplementedError(f"num_pos <= {lib.mml_crd_knn_max_positives()} (got {num_pos})")
    nbytes = lib.mml_crd_knn_workspace_bytes(n, B, D)
    if nbytes < 0:
        raise RuntimeError(f"knn_positives: unsupported sizes n={n} B={B} D={D}")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out_idx = torch.empty((B, num_pos), dtype=torch.int64, device=dev)
    out_sim = torch.empty((B, num_pos), dtype=torch.float32, device=dev)
    flags = torch.empty(B, dtype=torch.int32, device=dev) if return_flags else None
    _cabi.check(lib.mml_crd_knn_positives(
        _cabi.dptr(bank, torch.float32), n, D, _cabi.dptr(inv_norms, torch.float32) if inv_norms is not None else None,
        _cabi.dptr(row_labels, torch.int32), int(n_classes), _cabi.dptr(anchor_rows, torch.int64) if anchor_rows is not None else None,
        _cabi.dptr(queries, torch.float32) if queries is not None else None,
        _cabi.dptr(anchor_labels, torch.int64), B, int(num_pos), int(bool(exact_only)), _cabi.dptr(out_idx), _cabi.dptr(out_sim),
        _cabi.dptr(flags), _cabi.dptr(ws), ws.numel(), _cabi.cur_stream(dev)), "mml_crd_knn_positives")
    return (out_idx, out_sim, flags) if return_flags else (out_idx, out_sim)


def sharded_knn_positives(bank_local, row_begin, row_labels_local, anchor_rows, anchor_labels, num_pos, *, group=None,
                          n_classes=0, inv_norms=None):
    """`knn_positives` over a ROW-SHARDED bank ("over the sharded bank"): this rank holds rows
    [row_begin, row_begin + n_local) and `row_labels_local` for them; `anchor_rows` int64 [B_local] are GLOBAL row ids of this
    rank's anchors (same B_local on every rank), `anchor_labels` their classes.  -> (neighbours as GLOBAL row ids
    [B_local, P], similarities [B_local, P]), identical to the single-bank result.
    Exchange: anchors' ids / labels all-gathered (8 B each), their query rows supplied by whichever rank owns them (one
    all_reduce of [B_global, D]), every rank's exact local top-P per global anchor all-gathered (12 B per candidate), merged
    by the per-anchor sort kernel (score descending; ties: lower rank = lower rows first, rows ascending inside a rank).
    Bank rows never cross the wire."""
    import torch.distributed as dist
    from .crd_select import sort_columns
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local, D = bank_local.shape
    dev = bank_local.device
    Bl = anchor_rows.numel()
    rows_all = torch.empty(world * Bl, dtype=torch.int64, device=dev)
    labs_all = torch.empty(world * Bl, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(rows_all, anchor_rows.contiguous(), group=group)
    dist.all_gather_into_tensor(labs_all, anchor_labels.contiguous(), group=group)
    mine = (rows_all >= row_begin) & (rows_all < row_begin + n_local)
    Q = torch.zeros(world * Bl, D, dtype=torch.float32, device=dev)
    Q[mine] = bank_local.index_select(0, rows_all[mine] - row_begin)
    dist.all_reduce(Q, group=group)
    P_loc = min(int(num_pos), n_local)
    idx, sim = knn_positives(bank_local, row_labels_local, None, labs_all, P_loc, n_classes=n_classes, inv_norms=inv_norms, queries=Q)
    if P_loc < num_pos:                               # a shard smaller than num_pos: pad with candidates that never win
        pad = num_pos - P_loc
        idx = torch.cat((idx, idx.new_full((idx.shape[0], pad), -1)), 1)
        sim = torch.cat((sim, sim.new_full((sim.shape[0], pad), float("-inf"))), 1)
    idx = torch.where(idx >= 0, idx + row_begin, idx)
    P = int(num_pos)
    all_idx = torch.empty(world, world * Bl, P, dtype=torch.int64, device=dev)
    all_sim = torch.empty(world, world * Bl, P, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(all_idx, idx.contiguous(), group=group)
    dist.all_gather_into_tensor(all_sim, sim.contiguous(), group=group)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    cs = all_sim[:, sl].permute(1, 0, 2).reshape(Bl, world * P).contiguous()
    ci = all_idx[:, sl].permute(1, 0, 2).reshape(Bl, world * P).contiguous()
    order = sort_columns(cs, 0, world * P, descending=True, first=P)
    return ci.gather(1, order), cs.gather(1, order)


class _CenterBanks:
    """What `_ScoresFn` reads from a memory module, over the two [n_classes * (num_pos-1), D] centre tables."""

    def __init__(self, mem, c1, c2):
        self.memory_v1, self.memory_v2, self._T, self.params = c1, c2, mem._T, mem.params


class ContrastMemory(_crd.ContrastMemory):
    """
    memory buffer that supplies large amount of negative samples.
    return out_v1, out_v2: [batch size, K+num_pos, 1]                                    (CRD_criterion_v10.py:22-177)
    """

    def __init__(self, inputSize, outputSize, train_class_idx, K, T=0.07, momentum=0.5):
        nn.Module.__init__(self)
        self.nLem = outputSize
        self.unigrams = torch.ones(self.nLem)
        self.K = K
        self.class_idx = train_class_idx
        self.all_sample_labels = torch.zeros(outputSize)                                  # :34-38
        for c in range(len(self.class_idx)):
            self.all_sample_labels[torch.as_tensor(np.asarray(self.class_idx[c]), dtype=torch.long)] = c
        self.register_buffer('params', torch.tensor([K, T, -1, -1, momentum]))
        stdv = 1. / math.sqrt(inputSize / 3)
        self.register_buffer('memory_v1', torch.rand(outputSize, inputSize).mul_(2 * stdv).add_(-stdv))
        self.register_buffer('memory_v2', torch.rand(outputSize, inputSize).mul_(2 * stdv).add_(-stdv))
        self._refresh_scalars()
        self._pending = weakref.WeakSet()
        self._row_labels = None
        self._inv_norms = None           # [2, n] 1/|row| of both banks: full pass once, then only the batch's rows per step
        self._class_rows_dev = None      # crd_kmeans.ClassRows of `class_idx` on the banks' device ("centers")
        self.kmeans_generator = None     # torch.Generator of the k-means++ draws (None = the device's default generator)
        self.kmeans_init = None          # optional [2, n_classes, num_pos - 1, D]: fixed initial centres for banks 1 / 2

    def _apply(self, fn, *args, **kwargs):                  # no AliasMethod in this variant (idx is always supplied)
        out = nn.Module._apply(self, fn, *args, **kwargs)
        self._row_labels = None
        self._inv_norms = None
        self._class_rows_dev = None
        return out

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._inv_norms = None

    def invalidate_knn_cache(self):
        """Call after writing the banks other than through `forward` (the cached row norms would be stale)."""
        self._inv_norms = None

    def _norms(self):
        if self._inv_norms is None or self._inv_norms.device != self.memory_v1.device:
            self._inv_norms = torch.empty((2, self.memory_v1.shape[0]), dtype=torch.float32, device=self.memory_v1.device)
            knn_inv_norms(self.memory_v1, self._inv_norms[0])
            knn_inv_norms(self.memory_v2, self._inv_norms[1])
        return self._inv_norms

    def _update(self, v1, v2, y):
        super()._update(v1, v2, y)
        if self._inv_norms is not None:                      # the rows the momentum update rewrote
            knn_inv_norms(self.memory_v1, self._inv_norms[0], rows=y)
            knn_inv_norms(self.memory_v2, self._inv_norms[1], rows=y)

    def _labels_on(self, device):
        if self._row_labels is None or self._row_labels.device != device:
            self._row_labels = self.all_sample_labels.to(device=device, dtype=torch.int32)
        return self._row_labels

    def _class_rows(self, device):
        if self._class_rows_dev is None or self._class_rows_dev.device != device:
            self._class_rows_dev = _kmeans.ClassRows(self.class_idx, device)
        return self._class_rows_dev

    def _class_centers(self, bank, num_pos, which=0):
        """:84-92 -- [n_classes * (num_pos - 1), D]: the mean of every class's rows (num_pos == 2), or the num_pos - 1 k-means
        centres of every class (num_pos > 2; sklearn KMeans on the host in the reference, `crd_kmeans.class_kmeans` here,
        drawing its k-means++ initialisation from `self.kmeans_generator`, or started from `self.kmeans_init[which]`)."""
        cls = self._class_rows(bank.device)
        if num_pos == 2:
            if bank.shape[1] in (32, 64, 128, 256, 512):     # one pass over the bank: a single centre at 0 moves to the class mean
                mean = torch.zeros((cls.n_classes, 1, bank.shape[1]), dtype=torch.float32, device=bank.device)
                _kmeans.lloyd(bank.detach(), cls, mean)
                return mean.view(cls.n_classes, bank.shape[1])
            return torch.stack([bank.index_select(0, cls.rows[int(cls.offsets[c]):int(cls.offsets[c + 1])]).mean(0)
                                for c in range(cls.n_classes)]).contiguous()
        init = self.kmeans_init[which] if self.kmeans_init is not None else None
        centres = _kmeans.class_kmeans(bank, cls, num_pos - 1, init=init, generator=self.kmeans_generator)
        return centres.view(cls.n_classes * (num_pos - 1), bank.shape[1])

    def forward(self, num_pos, pos_extra, v1, v2, batch_label, y, idx=None):
        if not (v1.is_cuda and self.memory_v1.is_cuda):
            raise RuntimeError("ContrastMemory runs on CUDA tensors only (move the module with .cuda()/.to(device))")
        if idx is None:
            raise RuntimeError("CRD_criterion_v10.ContrastMemory has no sampler: contrast idx must be supplied (:64)")
        B = v1.size(0)
        K = self._K
        v1, v2, y = _crd._as_f32(v1), _crd._as_f32(v2), _crd._as_i64(y)
        idx = _crd._as_i64(idx).contiguous().view(B, K + 1)           # same RuntimeError as :65 for a wrong width
        batch_label = _crd._as_i64(batch_label)
        dev = v1.device
        if pos_extra == "neighbors":
            labels = self._labels_on(dev)
            anchors = idx[:, 0].contiguous()
            ncls = len(self.class_idx) if len(self.class_idx) <= 3 else 0
            inv = self._norms()
            nbr1, sim1 = knn_positives(self.memory_v1, labels, anchors, batch_label, num_pos, n_classes=ncls, inv_norms=inv[0])   # :69-76
            nbr2, sim2 = knn_positives(self.memory_v2, labels, anchors, batch_label, num_pos, n_classes=ncls, inv_norms=inv[1])   # :108-113
            # one gather pass over [neighbours of bank 1 | neighbours of bank 2 | the K negatives]; out_v2 reads bank 1 at
            # columns [0, P) + negatives, out_v1 reads bank 2 at columns [P, 2P) + negatives (:77-79, :114-119)
            P = num_pos
            cols = torch.cat((nbr1, nbr2, idx[:, 1:]), 1).contiguous()
            take1 = torch.cat((torch.arange(P, 2 * P, device=dev), torch.arange(2 * P, 2 * P + K, device=dev)))
            take2 = torch.cat((torch.arange(0, P, device=dev), torch.arange(2 * P, 2 * P + K, device=dev)))
        elif pos_extra == "centers":
            c1 = self._class_centers(self.memory_v1, num_pos, 0)
            c2 = self._class_centers(self.memory_v2, num_pos, 1)
            n_cls = len(self.class_idx)
            if n_cls != 3:
                raise RuntimeError("pos_extra='centers' hard-codes 3 classes in the reference (one_hot(..., num_classes=3), :60)")
            # rows of the centre tables read per sample: [own class's num_pos-1 centres | the other two classes'] (:60-61, :94-99)
            Q = num_pos - 1
            order = torch.tensor([[k] + [c for c in range(n_cls) if c != k] for k in range(n_cls)], device=dev)
            table = (order.view(n_cls, n_cls, 1) * Q + torch.arange(Q, device=dev)).view(n_cls, n_cls * Q)
            center_cols = table.index_select(0, batch_label).contiguous()
            cols = idx
        else:
            raise RuntimeError(f"pos_extra must be 'neighbors' or 'centers'; got {pos_extra!r}")   # reference: NameError later

        def raw_scores(mem, columns):
            r1, r2, _ = _crd.crd_scores(mem.memory_v1, mem.memory_v2, v1.detach(), v2.detach(), columns, self._T)
            return r1, r2

        if not self._z_ready:                                                                     # :121-129
            r1, r2 = raw_scores(self, cols)
            if pos_extra == "neighbors":
                r1, r2 = r1.index_select(1, take1), r2.index_select(1, take2)
            else:
                q1, q2 = raw_scores(_CenterBanks(self, c1, c2), center_cols)
                r1 = torch.cat((q1[:, :Q], r1, q1[:, Q:]), 1)
                r2 = torch.cat((q2[:, :Q], r2, q2[:, Q:]), 1)
            self.params[2] = r1.mean() * self.nLem
            self.params[3] = r2.mean() * self.nLem
            print("normalization constant Z_v1 is set to {:.1f}".format(self.params[2].item()))
            print("normalization constant Z_v2 is set to {:.1f}".format(self.params[3].item()))
            self._z_ready = True

        def scores(mem, columns, track):
            if torch.is_grad_enabled() and (v1.requires_grad or v2.requires_grad):
                undo = _crd._UndoLog()
                if track:
                    self._pending.add(undo)
                o1, o2 = _crd._ScoresFn.apply(v1, v2, mem, columns, undo)
                return o1.squeeze(2), o2.squeeze(2)
            o1, o2, _ = _crd.crd_scores(mem.memory_v1, mem.memory_v2, v1, v2, columns, self._T, Z=self.params[2:4])
            return o1, o2

        o1, o2 = scores(self, cols, True)
        if pos_extra == "neighbors":
            out_v1, out_v2 = o1.index_select(1, take1), o2.index_select(1, take2)
        else:
            q1, q2 = scores(_CenterBanks(self, c1, c2), center_cols, False)
            out_v1 = torch.cat((q1[:, :Q], o1, q1[:, Q:]), 1)      # [own centres | anchor + K negatives | other classes' centres]
            out_v2 = torch.cat((q2[:, :Q], o2, q2[:, Q:]), 1)
        out_v1, out_v2 = out_v1.unsqueeze(2).contiguous(), out_v2.unsqueeze(2).contiguous()
        self._update(v1, v2, y)                                                                   # :135-150
        if pos_extra == "neighbors":
            return out_v1, out_v2, sim1, sim2
        return out_v1, out_v2


class CRDLoss(nn.Module):
    """CRD Loss function (CRD_criterion_v10.py:181-232).

    Args: opt.s_dim / t_dim / feat_dim, opt.nce_k, opt.nce_t, opt.nce_m, opt.nce_p (number of positives),
    opt.pos_extra ("neighbors" | "centers"); n_data = bank rows; train_class_idx = per class, the dataset indices in it."""

    def __init__(self, opt, n_data, train_class_idx):
        super(CRDLoss, self).__init__()
        self.embed_s = Embed(opt.s_dim, opt.feat_dim)
        self.embed_t = Embed(opt.t_dim, opt.feat_dim)
        self.contrast = ContrastMemory(opt.feat_dim, n_data, train_class_idx, opt.nce_k, opt.nce_t, opt.nce_m)
        self.num_pos = opt.nce_p
        self.pos_extra = opt.pos_extra
        if self.pos_extra == "neighbors":
            self.criterion_t = ContrastLoss_v2(n_data)
            self.criterion_s = ContrastLoss_v2(n_data)
        else:
            self.criterion_t = ContrastLoss(n_data)
            self.criterion_s = ContrastLoss(n_data)

    def forward(self, sample_weights, f_s, f_t, batch_label, idx, contrast_idx=None):
        """-> (loss, per-sample loss [batch_size])"""
        f_s, f_t = _crd.embed_pair(self.embed_s, f_s, self.embed_t, f_t)      # the two heads on two streams
        if self.pos_extra == "neighbors":
            out_s, out_t, s_similarity, t_similarity = self.contrast(
                self.num_pos, self.pos_extra, f_s, f_t, batch_label, idx, contrast_idx)
            s_loss, s_sample_loss = self.criterion_s(sample_weights, out_s, self.num_pos, t_similarity)
            t_loss, t_sample_loss = self.criterion_t(sample_weights, out_t, self.num_pos, s_similarity)
        else:
            out_s, out_t = self.contrast(self.num_pos, self.pos_extra, f_s, f_t, batch_label, idx, contrast_idx)
            s_loss, s_sample_loss = self.criterion_s(sample_weights, out_s, self.num_pos)
            t_loss, t_sample_loss = self.criterion_t(sample_weights, out_t, self.num_pos)
        return s_loss + t_loss, s_sample_loss + t_sample_loss


def _log_terms(x, P, n_data):
    m = x.size(1) - P
    Pn = 1 / float(n_data)
    P_pos = x.narrow(1, 0, P)
    log_D1 = (P_pos / (P_pos + (m * Pn + eps))).log()
    P_neg = x.narrow(1, P, m)
    log_D0 = ((m * Pn) / (P_neg + (m * Pn + eps))).log()
    return log_D1, log_D0


class ContrastLoss(nn.Module):
    """class centres as the extra positives (CRD_criterion_v10.py:235-270)."""

    def __init__(self, n_data):
        super(ContrastLoss, self).__init__()
        self.n_data = n_data

    def forward(self, sample_weights, x, num_pos):
        P = num_pos
        bsz = x.shape[0]
        log_D1, log_D0 = _log_terms(x, P, self.n_data)
        if P > 1:
            sample_loss = -(log_D1.squeeze() + log_D0.sum(1).view(bsz, 1)).sum(1) / P           # :259
        else:
            sample_loss = -(log_D1.squeeze() + log_D0.sum(1).squeeze())                         # :262
        sample_loss = sample_weights.view(-1) * sample_loss
        return sample_loss.sum(0) / bsz, sample_loss


class ContrastLoss_v2(nn.Module):
    """KNN neighbours as positives, weighted by their similarity to the query (CRD_criterion_v10.py:274-311)."""

    def __init__(self, n_data):
        super(ContrastLoss_v2, self).__init__()
        self.n_data = n_data

    def forward(self, sample_weights, x, num_pos, knn_similarity):
        P = num_pos
        bsz = x.shape[0]
        log_D1, log_D0 = _log_terms(x, P, self.n_data)
        sample_loss = -((log_D1.squeeze() + log_D0.sum(1).view(bsz, 1)) * knn_similarity).sum(1) / knn_similarity.sum(1)   # :300-301
        sample_loss = sample_weights.view(-1) * sample_loss
        return sample_loss.sum(0) / bsz, sample_loss
