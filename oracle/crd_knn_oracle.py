"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's KNN / class-centre variant of CRD
(`MIA 2023/stage2_unimodal_student/CL_utils/CRD_criterion_v10.py`).  Only tests/, __graft_entry__.smoke() and bench.py's
CPU legs may import this module; the product path never does.

Pinned: `oracle/make_golden_knn.py` runs the UNMODIFIED reference classes on CPU (shims: `Tensor.cuda -> identity`, the
reference hard-codes `.cuda()` at :56-76) and `tests/test_oracle_golden.py` checks every function here against those
fixtures (`tests/golden/crdknn_*.npz`).  The k-means fits of pos_extra == "centers" with num_pos > 2 are random in the
reference (sklearn's k-means++ seeds itself); the generator replaces `KMeans` by a recording wrapper that draws the
initial centres itself, hands them to the real `sklearn.cluster.KMeans(init=...)` and stores them, so `kmeans_lloyd` here
is pinned against scikit-learn 1.9.0 on the reference's own call sites (`crdknn_kmeans_*`).
"""
from __future__ import annotations

import numpy as np
import torch

from . import crd_oracle as co

eps = 1e-7


def all_sample_labels(n, class_idx):
    """CRD_criterion_v10.py:34-38: class of every bank row (rows in no class keep 0)."""
    lab = torch.zeros(n)
    for c, rows in enumerate(class_idx):
        lab[torch.as_tensor(np.asarray(rows), dtype=torch.long)] = c
    return lab


def masked_cosine(memory, rows, labels, batch_label):
    """:69-70 -- class_mask[batch_label] * sklearn.cosine_similarity(memory[rows], memory): both sides L2-normalised
    (all-zero rows stay zero), rows of other classes score exactly 0.  -> [B, n]"""
    def normalize(x):
        nrm = x.norm(dim=1, keepdim=True)
        return x / torch.where(nrm == 0, torch.ones_like(nrm), nrm)
    sim = normalize(memory.index_select(0, rows)) @ normalize(memory).t()
    mask = (labels.view(1, -1) == batch_label.view(-1, 1).to(labels.dtype)).to(sim.dtype)
    return mask * sim


def knn_neighbors(memory, rows, labels, batch_label, num_pos):
    """:71-74 -- the num_pos best columns and their similarities, descending; ties to the smaller row (the reference's
    CUDA sort leaves them undefined)."""
    sim = masked_cosine(memory, rows, labels, batch_label)
    val, order = torch.sort(sim, dim=-1, descending=True, stable=True)
    return order[:, :num_pos], val[:, :num_pos]


def class_centers(memory, class_idx):
    """:84-88 with num_pos == 2: the mean row of every class.  -> [n_classes, D]"""
    return torch.stack([memory.index_select(0, torch.as_tensor(np.asarray(r), dtype=torch.long)).mean(0) for r in class_idx])


def kmeans_tolerance(X, tol=1e-4):
    """sklearn/cluster/_kmeans.py `_tolerance` (scikit-learn 1.9.0): tol * mean over features of the variance."""
    return float(np.mean(np.var(X, axis=0)) * tol)


def kmeans_lloyd(X, init, max_iter=300, tol=1e-4):
    """sklearn.cluster.KMeans(n_clusters=k, init=init, n_init=1).fit(X).cluster_centers_ restated (scikit-learn 1.9.0,
    `_kmeans_single_lloyd` + `lloyd_iter_chunked_dense`; scikit-learn is a dependency of the reference, not vendored in it --
    CRD_criterion_v10.py:6, :89-92): E-step argmin_j |c_j|^2 - 2 x.c_j (first minimum on ties), M-step mean of the
    members, stop when the labels repeat or the summed squared centre shift is <= the tolerance.  fp32 like the reference's
    bank.  An empty cluster keeps its centre (sklearn moves it to the farthest point; no fixture has one).
    -> (centres [k, D], iterations)"""
    X = np.asarray(X, dtype=np.float32)
    centres = np.array(init, dtype=np.float32, copy=True)
    limit = np.float32(kmeans_tolerance(X, tol))
    labels_old = None
    it = 0
    for it in range(1, max_iter + 1):
        scores = (centres * centres).sum(1, dtype=np.float32)[None, :] - np.float32(2) * (X @ centres.T)
        labels = scores.argmin(1)
        new = centres.copy()
        for j in range(centres.shape[0]):
            members = X[labels == j]
            if len(members):
                new[j] = members.mean(0, dtype=np.float32)
        shift = np.float32(((new - centres) ** 2).sum(dtype=np.float32))
        centres = new
        if labels_old is not None and np.array_equal(labels, labels_old):
            break
        if shift <= limit:
            break
        labels_old = labels
    return centres, it


def class_kmeans_centers(memory, class_idx, init):
    """:84-92 with num_pos > 2: k-means centres of every class's rows from the given initial centres [C, k, D] -> [C, k, D]."""
    out = []
    for c, r in enumerate(class_idx):
        X = memory.index_select(0, torch.as_tensor(np.asarray(r), dtype=torch.long)).numpy()
        out.append(torch.from_numpy(kmeans_lloyd(X, np.asarray(init[c]))[0]))
    return torch.stack(out)


def contrast_memory_v10_forward(memory_v1, memory_v2, params, class_idx, num_pos, pos_extra, v1, v2, batch_label, y, idx,
                                kmeans_init=None):
    """ContrastMemory.forward (:45-177).  Mutates params / banks; returns (out_v1, out_v2[, sim_v1, sim_v2]).
    kmeans_init [2, C, num_pos - 1, D]: the initial centres of the two banks' k-means fits (num_pos > 2; random in the reference)."""
    K, T = int(params[0].item()), params[1].item()
    n, D = memory_v1.shape
    B = v1.shape[0]
    labels = all_sample_labels(n, class_idx)
    outs, sims = [], []
    for which, (bank, v) in enumerate(((memory_v1, v2), (memory_v2, v1))):   # out_v2 from bank 1 (:107), out_v1 from bank 2 (:139)
        w = bank.index_select(0, idx.reshape(-1)).detach().view(B, K + 1, D)
        if pos_extra == "neighbors":
            nbr, sim = knn_neighbors(bank.detach(), idx[:, 0], labels, batch_label, num_pos)
            pos = bank.index_select(0, nbr.reshape(-1)).detach().view(B, num_pos, D)
            w = torch.cat((pos, w[:, 1:, :]), 1)                           # :79
            sims.append(sim)
        else:
            Q = num_pos - 1
            if num_pos == 2:
                cen = class_centers(bank.detach(), class_idx).view(-1, 1, D)       # [C, 1, D]
            else:
                cen = class_kmeans_centers(bank.detach(), class_idx, kmeans_init[which])   # [C, Q, D]
            n_cls = len(class_idx)
            others = torch.tensor([[c for c in range(n_cls) if c != k] for k in range(n_cls)])
            own = cen.index_select(0, batch_label).view(B, Q, D)
            neg = cen.index_select(0, others.index_select(0, batch_label).reshape(-1)).view(B, (n_cls - 1) * Q, D)
            w = torch.cat((own, w, neg), 1)                                # :98-104
        outs.append(torch.exp(torch.bmm(w, v.view(B, D, 1)) / T))
    out_v2, out_v1 = outs
    if params[2].item() < 0:
        params[2] = out_v1.mean().detach() * n
    if params[3].item() < 0:
        params[3] = out_v2.mean().detach() * n
    out_v1 = out_v1 / params[2].item()
    out_v2 = out_v2 / params[3].item()
    co.momentum_update_(memory_v1, y, v1.detach(), params[4].item())
    co.momentum_update_(memory_v2, y, v2.detach(), params[4].item())
    if pos_extra == "neighbors":
        return out_v1, out_v2, sims[0], sims[1]
    return out_v1, out_v2


def _log_terms(x, P, n_data):
    m = x.size(1) - P
    Pn = 1 / float(n_data)
    P_pos = x.narrow(1, 0, P)
    log_D1 = torch.div(P_pos, P_pos.add(m * Pn + eps)).log()
    P_neg = x.narrow(1, P, m)
    log_D0 = torch.div(P_neg.clone().fill_(m * Pn), P_neg.add(m * Pn + eps)).log()
    return log_D1, log_D0


def contrast_loss_centers(sample_weights, x, num_pos, n_data):
    """ContrastLoss.forward (:243-270) -> (loss, sample_loss)"""
    bsz, P = x.shape[0], num_pos
    log_D1, log_D0 = _log_terms(x, P, n_data)
    if P > 1:
        sample_loss = -((log_D1.squeeze() + log_D0.sum(1).view(bsz, 1).repeat(1, P))).sum(1) / P
    else:
        sample_loss = -(log_D1.squeeze() + log_D0.sum(1).squeeze())
    sample_loss = sample_weights.view(-1) * sample_loss
    return sample_loss.sum(0) / bsz, sample_loss


def contrast_loss_knn(sample_weights, x, num_pos, knn_similarity, n_data):
    """ContrastLoss_v2.forward (:282-311) -> (loss, sample_loss)"""
    bsz, P = x.shape[0], num_pos
    log_D1, log_D0 = _log_terms(x, P, n_data)
    sample_loss = -(torch.multiply(log_D1.squeeze() + log_D0.sum(1).view(bsz, 1).repeat(1, P), knn_similarity)).sum(1) \
        / knn_similarity.sum(1)
    sample_loss = sample_weights.view(-1) * sample_loss
    return sample_loss.sum(0) / bsz, sample_loss


def crd_loss_v10(sd, class_idx, num_pos, pos_extra, sample_weights, f_s, f_t, batch_label, idx, contrast_idx, n_data,
                 kmeans_init=None):
    """CRDLoss.forward (:208-232) over a state dict (mutated like the module's buffers) -> (loss, sample_loss, aux)."""
    v1 = co.embed_forward(f_s, sd, "embed_s.")
    v2 = co.embed_forward(f_t, sd, "embed_t.")
    res = contrast_memory_v10_forward(sd["contrast.memory_v1"], sd["contrast.memory_v2"], sd["contrast.params"], class_idx,
                                      num_pos, pos_extra, v1, v2, batch_label, idx, contrast_idx, kmeans_init)
    if pos_extra == "neighbors":
        out_s, out_t, s_sim, t_sim = res
        s_loss, s_sl = contrast_loss_knn(sample_weights, out_s, num_pos, t_sim, n_data)
        t_loss, t_sl = contrast_loss_knn(sample_weights, out_t, num_pos, s_sim, n_data)
    else:
        out_s, out_t = res
        s_loss, s_sl = contrast_loss_centers(sample_weights, out_s, num_pos, n_data)
        t_loss, t_sl = contrast_loss_centers(sample_weights, out_t, num_pos, n_data)
    return s_loss + t_loss, s_sl + t_sl, res
