"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's selection variant of CRD
(`MICCAI-2022/CL_utils/memory_new.py:225-397` ContrastMemory_v3, `CL_utils/CRD_loss.py:127-175`
5-arg CRDLoss, `CL_utils/CRD_loss.py:212-252` ContrastLoss_v2): the gather/exp/Z/update core of
`oracle/crd_oracle.py` over K+P columns plus cosine "relation" scores, two sorts and the pick of
P2 positives / K2 negatives.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module; the product path never does.

Also here: the re-weighted / one-directional variants of the MIA 2022 tree (`MIA 2022/CL_utils/memory_new.py:398-561`
ContrastMemory_v4, `:565-698` ContrastMemory_mono, `CL_utils/CRD_loss_v2.py:13-104`), pinned by
`oracle/make_golden_select_v4.py` -> `tests/golden/crdv4_*.npz`, `crdmono_*.npz`.

Pinned: `oracle/make_golden.py` runs the UNMODIFIED reference classes on CPU (the extra shim is
`Tensor.cuda -> identity`, the reference hard-codes `.cuda()` at memory_new.py:311-357) and
`tests/test_oracle_golden.py` checks every function here against those fixtures
(`tests/golden/crdsel_*.npz`).
"""
from __future__ import annotations

import numpy as np
import torch

from . import crd_oracle as co

eps = 1e-7      # CRD_loss.py:5


def make_params(K, P, T=0.07, momentum=0.5):
    """memory_new.py:244: params = [K, T, Z_v1, Z_v2, momentum, P] (fp32)."""
    return torch.tensor([K, T, -1, -1, momentum, P])


def relations(memory_v1, memory_v2, v1, v2, idx):
    """memory_new.py:288-292: cosine relation of every gathered row with the SAME-side embedding.
    t_relation pairs bank-1 rows with v1, s_relation pairs bank-2 rows with v2.  Each [B, K+P, 1]."""
    B, D = v1.shape
    cols = idx.shape[1]
    w1 = memory_v1.index_select(0, idx.reshape(-1)).view(B, cols, D)
    w2 = memory_v2.index_select(0, idx.reshape(-1)).view(B, cols, D)
    t_rel = torch.bmm(w1 / torch.norm(w1, dim=2, keepdim=True), (v1 / torch.norm(v1, dim=1, keepdim=True)).view(B, D, 1))
    s_rel = torch.bmm(w2 / torch.norm(w2, dim=2, keepdim=True), (v2 / torch.norm(v2, dim=1, keepdim=True)).view(B, D, 1))
    return t_rel, s_rel


def positive_picks(select_pos_mode, epoch, P, P2):
    """The numpy-RNG draw of memory_new.py:306-323 (positions in the sorted order); None for 'hard'.
    Consumes the GLOBAL numpy RNG exactly like the reference, so seeding numpy pins the picks."""
    if select_pos_mode == "hard":
        return None
    if select_pos_mode == "mid":
        return np.random.choice(np.arange(30, 100, 1), P2, replace=False)        # :311
    if select_pos_mode == "random":
        return np.random.randint(0, P, P2)                                         # :317
    if select_pos_mode == "curriculum":
        interval = 4 - np.ceil(3 * epoch)                                          # :320
        return np.random.randint(50 * (interval - 1), 50 * interval, P2)          # :321
    raise NotImplementedError(select_pos_mode)


def select_columns(diff, P, K, P2, K2, picks, select_neg_pairs="True"):
    """memory_new.py:298-361 on a precomputed diff = t_relation - s_relation [B, K+P]:
    -> int64 [B, P2 + (K2 | K)] column numbers into the K+P columns; column 0 is always the exact positive."""
    B = diff.shape[0]
    order = torch.sort(diff[:, :P], dim=1, descending=True)[1]                    # :303
    if picks is None:
        sel_pos = order[:, :P2].clone()                                            # :307
    else:
        sel_pos = order.index_select(1, torch.as_tensor(picks, dtype=torch.long))  # :314,318,322
    sel_pos[:, 0] = 0                                                              # :325
    if select_neg_pairs == "True":
        order_n = torch.sort(diff[:, P:P + K], dim=1, descending=False)[1]         # :342
        sel_neg = P + order_n[:, :K2]                                              # :345
    else:
        sel_neg = torch.arange(P, P + K).view(1, -1).repeat(B, 1)                  # :359-361
    return torch.cat((sel_pos, sel_neg), 1)


def contrast_memory_v3_forward(memory_v1, memory_v2, params, epoch, v1, v2, y, idx, *, P2, K2,
                               select_pos_mode="random", select_neg_pairs="True", picks="draw"):
    """ContrastMemory_v3.forward (memory_new.py:249-397) with a caller-supplied idx [B, K+P].
    Mutates `params` (first-call Z) and the banks in place; returns (out_v1, out_v2, sel) with
    out_* [B, P2+K2, 1] carrying autograd to v1 / v2 and sel the selected columns."""
    K, P = int(params[0].item()), int(params[5].item())
    T = params[1].item()
    n = memory_v1.size(0)
    raw_v1, raw_v2 = co.contrast_scores(memory_v1, memory_v2, params, v1, v2, idx)      # :270-278
    with torch.no_grad():
        t_rel, s_rel = relations(memory_v1, memory_v2, v1, v2, idx)
        diff = (t_rel - s_rel).squeeze(-1)
        if isinstance(picks, str):
            picks = positive_picks(select_pos_mode, epoch, P, P2)
        sel = select_columns(diff, P, K, P2, K2, picks, select_neg_pairs)
    out_v1 = raw_v1.squeeze(-1).gather(1, sel).unsqueeze(-1)                             # :333-334,350-351
    out_v2 = raw_v2.squeeze(-1).gather(1, sel).unsqueeze(-1)
    if params[2].item() < 0:                                                             # :367-374
        params[2] = out_v1.mean().detach() * n
    if params[3].item() < 0:
        params[3] = out_v2.mean().detach() * n
    out_v1 = torch.div(out_v1, params[2].item())
    out_v2 = torch.div(out_v2, params[3].item())
    co.momentum_update_(memory_v1, y, v1.detach(), params[4].item())                     # :381-395
    co.momentum_update_(memory_v2, y, v2.detach(), params[4].item())
    return out_v1, out_v2, sel


def contrast_loss_v2(x, P, n_data, sample_KD="False"):
    """ContrastLoss_v2.forward (CRD_loss.py:221-252): the first P columns of x [B, P+N, 1] are positives."""
    bsz = x.shape[0]
    N = x.size(1) - P
    m = N
    Pn = 1 / float(n_data)
    P_pos = x.narrow(1, 0, P)
    log_D1 = torch.div(P_pos, P_pos.add(m * Pn + eps)).log()
    P_neg = x.narrow(1, P, N)
    log_D0 = torch.div(P_neg.clone().fill_(m * Pn), P_neg.add(m * Pn + eps)).log()
    if sample_KD == "False":
        return -((log_D1.squeeze().sum(0) + log_D0.reshape(-1, 1).repeat(1, P).sum(0)) / bsz).sum(0) / P    # :241
    return -((log_D1.squeeze(-1) + (log_D0.repeat(1, 1, P)).sum(1))).sum(1) / P                            # :245


def crd_loss_v3(sd, epoch, f_s, f_t, idx, contrast_idx, n_data, *, P2, K2, select_pos_mode="random",
                select_neg_pairs="True", sample_KD="False", picks="draw"):
    """5-arg CRDLoss.forward (CRD_loss.py:153-175) over a state dict `sd` (mutated in place like the
    module's buffers).  The variant's Embed is a single Linear + L2 (CRD_loss.py:256-267)."""
    v1 = co.embed_forward(f_s, sd, "embed_s.")
    v2 = co.embed_forward(f_t, sd, "embed_t.")
    out_s, out_t, sel = contrast_memory_v3_forward(
        sd["contrast.memory_v1"], sd["contrast.memory_v2"], sd["contrast.params"], epoch, v1, v2, idx, contrast_idx,
        P2=P2, K2=K2, select_pos_mode=select_pos_mode, select_neg_pairs=select_neg_pairs, picks=picks)
    loss = contrast_loss_v2(out_s, P2, n_data, sample_KD) + contrast_loss_v2(out_t, P2, n_data, sample_KD)
    return loss, out_s, out_t, sel


def closed_form_multi_pos(rows1, rows2, v1, v2, T, Z1, Z2, n_data, P):
    """Loss and dL/dv of the sample_KD="False" criterion in closed form (the multi-positive analogue of
    SURVEY.md A.3), in float64.  rows1/rows2 [B, P+N, D] are the SELECTED rows of bank 1 / bank 2."""
    B, cols, D = rows1.shape
    N = cols - P
    c = N / float(n_data) + eps
    r1, r2, a, b = rows1.double(), rows2.double(), v1.double(), v2.double()
    loss = 0.0
    grads = []
    for rows, v, Z in ((r2, a, Z1), (r1, b, Z2)):          # out_v1 = bank2 . v1 ; out_v2 = bank1 . v2
        x = torch.exp(torch.einsum("bkd,bd->bk", rows, v) / T) / Z
        pos, neg = x[:, :P], x[:, P:]
        loss = loss - ((pos / (pos + c)).log().sum() / P + ((N / float(n_data)) / (neg + c)).log().sum()) / B
        g = torch.empty_like(x)
        g[:, :P] = -(c / (pos + c)) / (T * B * P)
        g[:, P:] = (neg / (neg + c)) / (T * B)
        grads.append(torch.einsum("bk,bkd->bd", g, rows))
    return loss, grads[0], grads[1]


# ---- MIA 2022 tree: ContrastMemory_v4 / ContrastMemory_mono (memory_new.py:398-698), CRD_loss_v2.py ----
def positive_picks_mono(select_pos_mode, epoch, P, P2):
    """memory_new.py:639-651: as `positive_picks`, except 'mid' draws randint(50, 100) (:643)."""
    if select_pos_mode == "mid":
        return np.random.randint(50, 100, P2)
    return positive_picks(select_pos_mode, epoch, P, P2)


def _pick_positive_columns(gap_pos, picks, P2):
    """Sort `gap_pos` [B, P] descending and take the first P2 / the picked ranks; column 0 is forced to the anchor."""
    order = torch.sort(gap_pos, dim=1, descending=True)[1]
    sel = order[:, :P2].clone() if picks is None else order.index_select(1, torch.as_tensor(picks, dtype=torch.long))
    sel[:, 0] = 0
    return sel


def contrast_memory_v4_forward(memory_v1, memory_v2, params, epoch, v1, v2, y, idx, *, P2, select_pos_mode="mid",
                               neg_reweight="True", picks="draw"):
    """ContrastMemory_v4.forward (memory_new.py:422-561).  Relations (:462-466) pair bank-1 rows with v1 (the STUDENT side
    here) and bank-2 rows with v2 (teacher); positives by descending t - s (:476), every negative kept and, when
    neg_reweight == "True", scaled by s - t + 1 (:493-499).  Z is the mean of the combined scores (:512-519).
    Mutates params / banks; returns (out_v1, out_v2, sel_pos) with out_* [B, P2+K, 1]."""
    K, P = int(params[0].item()), int(params[5].item())
    n = memory_v1.size(0)
    raw_v1, raw_v2 = co.contrast_scores(memory_v1, memory_v2, params, v1, v2, idx)      # :441-450
    with torch.no_grad():
        s_rel, t_rel = relations(memory_v1, memory_v2, v1, v2, idx)                     # roles swapped w.r.t. v3
        if isinstance(picks, str):
            picks = positive_picks(select_pos_mode, epoch, P, P2)
        sel_pos = _pick_positive_columns((t_rel - s_rel).squeeze(-1)[:, :P], picks, P2)
        weight = (s_rel - t_rel + 1).narrow(1, P, K)                                    # :497
    outs = []
    for raw in (raw_v1, raw_v2):
        pos = raw.squeeze(-1).gather(1, sel_pos).unsqueeze(-1)                          # :489-490
        neg = raw.narrow(1, P, K)
        if neg_reweight == "True":
            neg = neg * weight                                                          # :498-499
        elif neg_reweight != "False":
            raise RuntimeError("neg_reweight must be 'True' or 'False'")                # reference: UnboundLocalError
        outs.append(torch.cat((pos, neg), 1))                                           # :508-509
    out_v1, out_v2 = outs
    if params[2].item() < 0:
        params[2] = out_v1.mean().detach() * n
    if params[3].item() < 0:
        params[3] = out_v2.mean().detach() * n
    out_v1 = torch.div(out_v1, params[2].item())
    out_v2 = torch.div(out_v2, params[3].item())
    co.momentum_update_(memory_v1, y, v1.detach(), params[4].item())
    co.momentum_update_(memory_v2, y, v2.detach(), params[4].item())
    return out_v1, out_v2, sel_pos


def contrast_memory_mono_forward(memory_v1, memory_v2, params, epoch, v1, v2, y, idx, *, P2, select_pos_mode="hard",
                                 picks="draw"):
    """ContrastMemory_mono.forward (memory_new.py:589-698): params = [P, K, T, Z_v2, momentum] (:586); v1 = teacher,
    v2 = student.  out_v2 = exp(bank1 . v2 / T) (:608-611); positives by descending t - s with t = (bank 1, v1),
    s = (bank 2, v2) (:625-635); all K negatives (:663).  Returns (out_v2 [B, P2+K, 1], sel_pos); both banks updated."""
    P, K = int(params[0].item()), int(params[1].item())
    T = params[2].item()
    n = memory_v1.size(0)
    B, D = v2.shape
    rows1 = memory_v1.index_select(0, idx.reshape(-1)).detach().view(B, K + P, D)
    raw_v2 = torch.exp(torch.bmm(rows1, v2.view(B, D, 1)).div(T))
    with torch.no_grad():
        t_rel, s_rel = relations(memory_v1, memory_v2, v1, v2, idx)
        if isinstance(picks, str):
            picks = positive_picks_mono(select_pos_mode, epoch, P, P2)
        sel_pos = _pick_positive_columns((t_rel - s_rel).squeeze(-1)[:, :P], picks, P2)
    out_v2 = torch.cat((raw_v2.squeeze(-1).gather(1, sel_pos).unsqueeze(-1), raw_v2.narrow(1, P, K)), 1)
    if params[3].item() < 0:                                                             # :668-671
        params[3] = out_v2.mean().detach() * n
    out_v2 = torch.div(out_v2, params[3].item())
    co.momentum_update_(memory_v1, y, v1.detach(), params[4].item())
    co.momentum_update_(memory_v2, y, v2.detach(), params[4].item())
    return out_v2, sel_pos


def crd_loss_v4(sd, epoch, f_s, f_t, idx, contrast_idx, n_data, *, P2, select_pos_mode, neg_reweight="True",
                sample_KD="False", picks="draw"):
    """`CRD_loss_v2.py:13-55` CRDLoss.forward over a state dict (mutated like the module's buffers)."""
    v1 = co.embed_forward(f_s, sd, "embed_s.")
    v2 = co.embed_forward(f_t, sd, "embed_t.")
    out_s, out_t, sel_pos = contrast_memory_v4_forward(
        sd["contrast.memory_v1"], sd["contrast.memory_v2"], sd["contrast.params"], epoch, v1, v2, idx, contrast_idx,
        P2=P2, select_pos_mode=select_pos_mode, neg_reweight=neg_reweight, picks=picks)
    loss = contrast_loss_v2(out_s, P2, n_data, sample_KD) + contrast_loss_v2(out_t, P2, n_data, sample_KD)
    return loss, out_s, out_t, sel_pos


def crd_loss_mono(sd, epoch, f_s, f_t, idx, contrast_idx, n_data, *, P2, select_pos_mode, sample_KD="False", picks="draw"):
    """`CRD_loss_v2.py:58-104` CRDLoss_v2.forward: teacher feature detached + L2-normalised (:92,95), student embedded (:93),
    the memory called as (epoch, f_t, f_s, ...) (:98)."""
    f_t = co.l2_normalize(f_t.clone().detach(), 2)
    v_s = co.embed_forward(f_s, sd, "embed_s.")
    out_s, sel_pos = contrast_memory_mono_forward(
        sd["contrast.memory_v1"], sd["contrast.memory_v2"], sd["contrast.params"], epoch, f_t, v_s, idx, contrast_idx,
        P2=P2, select_pos_mode=select_pos_mode, picks=picks)
    return contrast_loss_v2(out_s, P2, n_data, sample_KD), out_s, sel_pos
