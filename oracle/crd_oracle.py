"""CPU oracle for the CRD contrastive-distillation path.  TEST INFRASTRUCTURE ONLY.

This is a functional restatement, in plain torch-CPU / numpy arithmetic, of the
reference's `MICCAI-2022/CL_utils/CRD_criterion.py`.  It is the checker for the
CUDA path; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  The product package
never does (and has no CPU fallback).

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference
modules from /root/reference (with the three import shims of SURVEY.md §8c) on
seeded inputs and stores inputs+outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks every function here against those
fixtures (bit-exact for alias tables / draws / indices, <=1e-6 rel for floats).

All `file:line` citations are into /root/reference/MICCAI-2022/.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

NCE_EPS = 1e-7  # CL_utils/CRD_criterion.py:5


# --------------------------------------------------------------------------- #
# Normalize / Embed                                   CRD_criterion.py:219-245
# --------------------------------------------------------------------------- #
def l2_normalize(x: torch.Tensor, power: float = 2) -> torch.Tensor:
    """CRD_criterion.py:242-245 -- divide by the raw p-norm, no epsilon."""
    nrm = x.pow(power).sum(1, keepdim=True).pow(1.0 / power)
    return x.div(nrm)


def embed_forward(x: torch.Tensor, sd: dict, prefix: str) -> torch.Tensor:
    """CRD_criterion.py:229-233.  `sd` holds `<prefix>linear.0.*`/`linear.2.*`
    (Linear-ReLU-Linear head) or `<prefix>linear.*` (single-Linear variant,
    `MIA 2022/CL_utils/CRD_criterion.py:223`)."""
    h = x.reshape(x.shape[0], -1)
    if prefix + "linear.weight" in sd:
        h = F.linear(h, sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])
    else:
        h = F.linear(h, sd[prefix + "linear.0.weight"], sd[prefix + "linear.0.bias"])
        h = torch.relu(h)
        h = F.linear(h, sd[prefix + "linear.2.weight"], sd[prefix + "linear.2.bias"])
    return l2_normalize(h, 2)


# --------------------------------------------------------------------------- #
# ContrastMemory.forward                               CRD_criterion.py:25-81
# --------------------------------------------------------------------------- #
def memory_init(n_rows: int, dim: int, generator: torch.Generator | None = None):
    """CRD_criterion.py:21-23 -- both banks ~ U(-stdv, stdv), stdv = 1/sqrt(D/3),
    bank 1 drawn first then bank 2."""
    stdv = 1.0 / math.sqrt(dim / 3)
    m1 = torch.rand(n_rows, dim, generator=generator).mul_(2 * stdv).add_(-stdv)
    m2 = torch.rand(n_rows, dim, generator=generator).mul_(2 * stdv).add_(-stdv)
    return m1, m2


def make_params(K: int, T: float = 0.07, momentum: float = 0.5) -> torch.Tensor:
    """CRD_criterion.py:20 -- fp32 buffer [K, T, Z_v1, Z_v2, momentum], Z unset = -1."""
    return torch.tensor([K, T, -1, -1, momentum])


def contrast_scores(memory_v1, memory_v2, params, v1, v2, idx):
    """CRD_criterion.py:41-49: raw exp(dot/T) scores before the Z division.
    Returns (raw_v1, raw_v2), each [B, K+1, 1].  raw_v1 pairs v1 with bank 2."""
    T = params[1].item()  # fp32(0.07) widened to a python double, :27
    B, D = v1.shape
    cols = idx.shape[1]
    rows1 = memory_v1.index_select(0, idx.reshape(-1)).detach().view(B, cols, D)
    raw_v2 = torch.exp(torch.bmm(rows1, v2.view(B, D, 1)).div(T))
    rows2 = memory_v2.index_select(0, idx.reshape(-1)).detach().view(B, cols, D)
    raw_v1 = torch.exp(torch.bmm(rows2, v1.view(B, D, 1)).div(T))
    return raw_v1, raw_v2


def momentum_update_(bank, rows, v, momentum):
    """CRD_criterion.py:66-72 (bank 1) / :74-79 (bank 2): in-place row update
    r <- normalize(m*r + (1-m)*v) using the pre-step row values."""
    with torch.no_grad():
        pos = bank.index_select(0, rows.view(-1))
        pos.mul_(momentum)
        pos.add_(torch.mul(v, 1 - momentum))
        nrm = pos.pow(2).sum(1, keepdim=True).pow(0.5)
        bank.index_copy_(0, rows, pos.div(nrm))


def contrast_memory_forward(memory_v1, memory_v2, params, v1, v2, y, idx):
    """ContrastMemory.forward with a caller-supplied idx (CRD_criterion.py:25-81).
    Mutates `params[2:4]` on the first call and both banks on every call.
    Returns (out_v1, out_v2), each [B, K+1, 1]."""
    K = int(params[0].item())
    assert idx.shape[1] == K + 1, "idx must have K+1 columns (:42)"
    momentum = params[4].item()
    n_rows = memory_v1.size(0)
    raw_v1, raw_v2 = contrast_scores(memory_v1, memory_v2, params, v1, v2, idx)
    if params[2].item() < 0:      # :52-55
        params[2] = raw_v1.mean() * n_rows
    if params[3].item() < 0:      # :56-59
        params[3] = raw_v2.mean() * n_rows
    Z_v1 = params[2].clone().detach().item()
    Z_v2 = params[3].clone().detach().item()
    out_v1 = raw_v1.div(Z_v1).contiguous()   # :62
    out_v2 = raw_v2.div(Z_v2).contiguous()   # :63
    momentum_update_(memory_v1, y, v1, momentum)
    momentum_update_(memory_v2, y, v2, momentum)
    return out_v1, out_v2


# --------------------------------------------------------------------------- #
# ContrastLoss (NCE criterion, Eq. 18)               CRD_criterion.py:199-216
# --------------------------------------------------------------------------- #
def nce_loss(x: torch.Tensor, n_data: int) -> torch.Tensor:
    """x: [B, K+1, 1], column 0 = positive.  Returns shape [1]."""
    B = x.shape[0]
    m = x.size(1) - 1
    Pn = 1 / float(n_data)
    pos = x.select(1, 0)
    log_d1 = torch.div(pos, pos.add(m * Pn + NCE_EPS)).log_()
    neg = x.narrow(1, 1, m)
    log_d0 = torch.div(torch.full_like(neg, m * Pn), neg.add(m * Pn + NCE_EPS)).log_()
    return -(log_d1.sum(0) + log_d0.view(-1, 1).sum(0)) / B


def crd_loss(sd: dict, f_s, f_t, idx, contrast_idx, n_data: int):
    """CRDLoss.forward (CRD_criterion.py:167-188) over a state_dict `sd` with the
    reference's key names (`embed_s.*`, `embed_t.*`, `contrast.params`,
    `contrast.memory_v1`, `contrast.memory_v2`).  Mutates the contrast buffers.
    Returns (loss[1], v_s, v_t) -- the embeddings are returned for inspection."""
    v_s = embed_forward(f_s, sd, "embed_s.")
    v_t = embed_forward(f_t, sd, "embed_t.")
    out_s, out_t = contrast_memory_forward(
        sd["contrast.memory_v1"], sd["contrast.memory_v2"], sd["contrast.params"],
        v_s, v_t, idx, contrast_idx)
    return nce_loss(out_s, n_data) + nce_loss(out_t, n_data), v_s, v_t


def crd_closed_form(memory_v1, memory_v2, v1, v2, idx, T, Z1, Z2, n_data):
    """Second oracle: loss and dL/dv1, dL/dv2 in closed
    form, float64, from the pre-update banks.  Used to cross-check the fused
    kernel's gradient without autograd."""
    B, D = v1.shape
    K = idx.shape[1] - 1
    m1 = memory_v1.double()[idx]          # [B,K+1,D]
    m2 = memory_v2.double()[idx]
    c = K * (1 / float(n_data)) + NCE_EPS
    kp = K * (1 / float(n_data))

    def one_side(rows, v, Z):
        x = torch.exp((rows @ v.double().unsqueeze(2)).squeeze(2) / T) / Z   # [B,K+1]
        loss = -(torch.log(x[:, 0] / (x[:, 0] + c)).sum()
                 + torch.log(kp / (x[:, 1:] + c)).sum()) / B
        dx = torch.empty_like(x)
        dx[:, 0] = -(1 / x[:, 0] - 1 / (x[:, 0] + c)) / B
        dx[:, 1:] = 1 / (x[:, 1:] + c) / B
        g = dx * x / T
        return loss, (g.unsqueeze(2) * rows).sum(1), x
    l1, gv1, x1 = one_side(m2, v1, Z1)    # out_v1: v1 against bank 2
    l2, gv2, x2 = one_side(m1, v2, Z2)    # out_v2: v2 against bank 1
    return l1 + l2, gv1, gv2, x1, x2


# --------------------------------------------------------------------------- #
# AliasMethod                                        CRD_criterion.py:84-141
# --------------------------------------------------------------------------- #
def alias_build(probs: np.ndarray):
    """Vose tables with the reference's LIFO pairing (CRD_criterion.py:88-123).
    `probs`: fp32 [n], ALREADY normalised the way the reference does it
    (`probs.div_(probs.sum())` when the sum exceeds 1, :90-91 -- that torch call
    stays on the host side).  Returns (prob fp32[n], alias int64[n]).
    Every arithmetic step is fp32, as it is on the reference's fp32 tensors."""
    p = np.asarray(probs, dtype=np.float32)
    n = p.shape[0]
    prob = np.zeros(n, dtype=np.float32)
    alias = np.zeros(n, dtype=np.int64)
    kf = np.float32(n)                       # python int * fp32 0-dim tensor -> fp32
    prob[:] = kf * p                         # :101
    small_mask = prob < np.float32(1.0)      # :102
    smaller = np.nonzero(small_mask)[0].tolist()
    larger = np.nonzero(~small_mask)[0].tolist()
    one = np.float32(1.0)
    while smaller and larger:                # :110-120
        s = smaller.pop()
        l = larger.pop()
        alias[s] = l
        prob[l] = np.float32(np.float32(prob[l] - one) + prob[s])
        if prob[l] < one:
            smaller.append(l)
        else:
            larger.append(l)
    for j in smaller + larger:               # :122-123
        prob[j] = one
    return prob, alias


def alias_select(kk: np.ndarray, b: np.ndarray, alias: np.ndarray) -> np.ndarray:
    """CRD_criterion.py:138-141 given the raw draws: kk ~ U{0..n-1} (:133) and
    b = bernoulli(prob[kk]) (:137).  Returns kk*b + alias[kk]*(1-b), int64."""
    bl = b.astype(np.int64)
    return kk * bl + alias[kk] * (1 - bl)
