"""CRD contrastive distiller -- host-side mirror of the reference's
`MICCAI-2022/CL_utils/CRD_criterion.py` (same class names, constructor and
`forward` signatures, buffer/parameter names, error behaviour) over the
hand-written sm_100a kernels in libmml_b200.so.

    CRDLoss(opt).forward(f_s, f_t, idx, contrast_idx=None) -> Tensor[1]      (:158,167)
    ContrastMemory(inputSize, outputSize, K, T, momentum).forward(v1, v2, y, idx=None)
        -> (out_v1, out_v2), each [B, K+1, 1]                                 (:12,25,81)
    AliasMethod(probs) / .cuda() / .draw(N)                                   (:88,125,129)
    ContrastLoss(n_data).forward(x) -> Tensor[1]                              (:195,199)
    Embed(dim_in, dim_out), Normalize(power)                                  (:221,238)

What differs from the reference, on purpose:
  * `CRDLoss.forward` never materialises the gathered rows [B, K+1, D] nor the
    scores: one fused kernel yields the loss and dL/dv (stashed for backward).
  * no `.item()` host syncs per call: K, T, momentum are cached python scalars;
    Z stays on the device (one sync on the very first call, where the reference
    prints Z).
  * CUDA only.  There is no CPU path; the oracle under `oracle/` is test-only.
"""
from __future__ import annotations

import math
import os
import weakref

import torch
from torch import nn
from torch.autograd.function import once_differentiable

from . import _cabi

eps = 1e-7   # CRD_criterion.py:5

# Optional profiling hook (bench.py): an object with .start(name, stream_device) / .stop(name) that records
# CUDA events on the launching stream around the gather launch.  None in normal use.
KERNEL_TIMER = None


# --------------------------------------------------------------------------- #
# thin typed wrappers around the C ABI
# --------------------------------------------------------------------------- #
def _workspace(B: int, cols: int, D: int, device) -> torch.Tensor:
    n = _cabi.lib().mml_crd_workspace_bytes(B, cols, D)
    return torch.empty(n, dtype=torch.uint8, device=device)


def _as_f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError(f"CRD kernels compute in fp32; got {t.dtype}")
    return t.contiguous()


def _idx_arg(idx: torch.Tensor):
    """(device pointer, element size) of an int64 / int32 index tensor."""
    if idx.dtype not in (torch.int64, torch.int32):
        raise RuntimeError(f"index tensors must be int64 or int32; got {idx.dtype}")
    return _cabi.dptr(idx), idx.element_size()


def _as_i64(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int64:
        raise RuntimeError(f"index tensors must be int64 (LongTensor); got {t.dtype}")
    return t.contiguous()


def crd_fused_loss_grad(bank1, bank2, v1, v2, idx, T, Z, n_data, nce_k, *, seg_ptr=None, pos_flag=None,
                        batch_norm=None, want_out=False, want_sums=False, cols=None):
    """-> (loss[1] | None, sums[4] | None, g1[B,D], g2[B,D], out_v1 | None, out_v2 | None)"""
    B, D = v1.shape
    cols = idx.shape[1] if seg_ptr is None else int(cols if cols is not None else nce_k + 1)
    dev = v1.device
    ws = _workspace(B, cols, D, dev)
    g1 = torch.empty_like(v1)
    g2 = torch.empty_like(v2)
    loss = torch.empty(1, dtype=torch.float32, device=dev) if not want_sums else None
    sums = torch.empty(4, dtype=torch.float32, device=dev) if want_sums else None
    out1 = torch.empty(idx.shape, dtype=torch.float32, device=dev) if want_out else None
    out2 = torch.empty(idx.shape, dtype=torch.float32, device=dev) if want_out else None
    if KERNEL_TIMER is not None:
        KERNEL_TIMER.start("crd_fused_loss_grad", dev)
    rc = _cabi.lib().mml_crd_fused_loss_grad(
        _cabi.dptr(bank1), _cabi.dptr(bank2), bank1.shape[0], D, _cabi.dptr(v1), _cabi.dptr(v2),
        *_idx_arg(idx), _cabi.dptr(seg_ptr), _cabi.dptr(pos_flag), B, cols,
        float(T), _cabi.dptr(Z), int(n_data), int(nce_k), int(batch_norm if batch_norm is not None else B),
        _cabi.dptr(loss), _cabi.dptr(sums), _cabi.dptr(g1), _cabi.dptr(g2), _cabi.dptr(out1), _cabi.dptr(out2),
        _cabi.dptr(ws), ws.numel(), _cabi.cur_stream(dev))
    if KERNEL_TIMER is not None:
        KERNEL_TIMER.stop("crd_fused_loss_grad", dev)
    _cabi.check(rc, "mml_crd_fused_loss_grad")
    return loss, sums, g1, g2, out1, out2


def crd_scores(bank1, bank2, v1, v2, idx, T, *, Z=None, set_Z=None, seg_ptr=None, cols=None,
               want_out=True, want_sums=False):
    """-> (out_v1 | None, out_v2 | None, sums[4] | None)"""
    B, D = v1.shape
    cols = idx.shape[1] if seg_ptr is None else int(cols)
    dev = v1.device
    ws = _workspace(B, cols, D, dev)
    out1 = torch.empty(idx.shape, dtype=torch.float32, device=dev) if want_out else None
    out2 = torch.empty(idx.shape, dtype=torch.float32, device=dev) if want_out else None
    sums = torch.empty(4, dtype=torch.float32, device=dev) if want_sums else None
    rc = _cabi.lib().mml_crd_scores(
        _cabi.dptr(bank1), _cabi.dptr(bank2), bank1.shape[0], D, _cabi.dptr(v1), _cabi.dptr(v2),
        *_idx_arg(idx), _cabi.dptr(seg_ptr), B, cols, float(T), _cabi.dptr(Z),
        _cabi.dptr(sums), _cabi.dptr(set_Z), _cabi.dptr(out1), _cabi.dptr(out2),
        _cabi.dptr(ws), ws.numel(), _cabi.cur_stream(dev))
    _cabi.check(rc, "mml_crd_scores")
    return out1, out2, sums


def crd_weighted_rows(bank1, bank2, idx, coef1, coef2, *, seg_ptr=None, cols=None, B=None):
    """g1[b] = sum_k coef1[b,k] bank2[idx[b,k]],  g2[b] = sum_k coef2[b,k] bank1[idx[b,k]]."""
    D = bank1.shape[1]
    if seg_ptr is None:
        B, cols = idx.shape
    dev = bank1.device
    ws = _workspace(B, cols, D, dev)
    g1 = torch.empty(B, D, dtype=torch.float32, device=dev)
    g2 = torch.empty(B, D, dtype=torch.float32, device=dev)
    rc = _cabi.lib().mml_crd_weighted_rows(
        _cabi.dptr(bank1), _cabi.dptr(bank2), bank1.shape[0], D, *_idx_arg(idx),
        _cabi.dptr(seg_ptr), _cabi.dptr(coef1), _cabi.dptr(coef2), B, cols, _cabi.dptr(g1), _cabi.dptr(g2),
        _cabi.dptr(ws), ws.numel(), _cabi.cur_stream(dev))
    _cabi.check(rc, "mml_crd_weighted_rows")
    return g1, g2


def crd_memory_update(bank1, bank2, v1, v2, y, momentum, row_begin=0, row_end=None):
    """In-place momentum + L2-renorm update of rows y of both banks (CRD_criterion.py:66-79)."""
    B, D = v1.shape
    if row_end is None:
        row_end = row_begin + bank1.shape[0]
    rc = _cabi.lib().mml_crd_memory_update(
        _cabi.dptr(bank1), _cabi.dptr(bank2), D, _cabi.dptr(v1), _cabi.dptr(v2), _cabi.dptr(y, torch.int64), B,
        float(momentum), int(row_begin), int(row_end), _cabi.cur_stream(v1.device))
    _cabi.check(rc, "mml_crd_memory_update")


# --------------------------------------------------------------------------- #
# AliasMethod                                         CRD_criterion.py:84-141
# --------------------------------------------------------------------------- #
class AliasMethod(object):
    """Vose alias sampler.  Tables are built on the host by the C library
    (`mml_alias_build_host`, bit-exact with the reference's python loop, without
    its 33 us/row cost); `draw` keeps the reference's two torch RNG calls
    (`random_`, `bernoulli`) and fuses the table lookups/select around them."""

    def __init__(self, probs):
        if probs.sum() > 1:                      # :90-91 (mutates the caller's tensor, as the reference does)
            probs.div_(probs.sum())
        K = len(probs)
        p = probs.detach().to("cpu", torch.float32).contiguous()
        self.prob = torch.zeros(K)
        self.alias = torch.zeros(K, dtype=torch.long)
        rc = _cabi.lib().mml_alias_build_host(_cabi.hptr(p), K, _cabi.hptr(self.prob), _cabi.hptr(self.alias))
        _cabi.check(rc, "mml_alias_build_host")

    def cuda(self, device=None):
        self.prob = self.prob.cuda(device)
        self.alias = self.alias.cuda(device)

    def to(self, device):
        self.prob = self.prob.to(device)
        self.alias = self.alias.to(device)

    def draw(self, N, y=None, cols=None):
        """Draw N samples from the multinomial (:129-141).  With `y`/`cols` the
        ContrastMemory caller's column-0 overwrite (:39) is fused into the select."""
        if not self.prob.is_cuda:
            raise RuntimeError("AliasMethod.draw runs on the GPU only: call .cuda() first (no CPU fallback)")
        K = self.alias.size(0)
        dev = self.prob.device
        kk = torch.zeros(N, dtype=torch.long, device=dev).random_(0, K)          # :133
        p = torch.empty(N, dtype=torch.float32, device=dev)
        st = _cabi.cur_stream(dev)
        _cabi.check(_cabi.lib().mml_alias_gather_prob(_cabi.dptr(self.prob), _cabi.dptr(kk), N, _cabi.dptr(p), st),
                    "mml_alias_gather_prob")
        b = torch.bernoulli(p)                                                    # :137
        out = torch.empty(N, dtype=torch.long, device=dev)
        _cabi.check(_cabi.lib().mml_alias_select(
            _cabi.dptr(self.alias), _cabi.dptr(kk), _cabi.dptr(b), N,
            _cabi.dptr(y, torch.int64) if y is not None else None, int(cols or 1), _cabi.dptr(out), st),
            "mml_alias_select")
        return out


# --------------------------------------------------------------------------- #
# ContrastMemory                                        CRD_criterion.py:8-81
# --------------------------------------------------------------------------- #
class _UndoLog:
    """Pre-update copies of bank rows touched after a differentiable
    `ContrastMemory.forward`, so its backward sees the rows it scored."""

    def __init__(self):
        self.entries = []      # (y, old_rows_bank1, old_rows_bank2), oldest first


class _ScoresFn(torch.autograd.Function):
    """out_v1/out_v2 of ContrastMemory.forward with gradients to v1/v2."""

    @staticmethod
    def forward(ctx, v1, v2, mem, idx, undo):
        out1, out2, _ = crd_scores(mem.memory_v1, mem.memory_v2, v1, v2, idx, mem._T, Z=mem.params[2:4])
        ctx.mem, ctx.undo, ctx.T = mem, undo, mem._T
        out1, out2 = out1.unsqueeze(2), out2.unsqueeze(2)
        ctx.save_for_backward(idx, out1, out2)
        return out1, out2

    @staticmethod
    @once_differentiable
    def backward(ctx, go1, go2):
        idx, out1, out2 = ctx.saved_tensors
        mem = ctx.mem
        B, cols = idx.shape
        coef1 = (go1 * out1 / ctx.T).reshape(B, cols).contiguous()     # d out/d dot = out / T
        coef2 = (go2 * out2 / ctx.T).reshape(B, cols).contiguous()
        g1, g2 = crd_weighted_rows(mem.memory_v1, mem.memory_v2, idx, coef1, coef2)
        if ctx.undo.entries:      # rows updated since the forward: swap in their pre-update values
            ys = torch.cat([e[0] for e in ctx.undo.entries])
            o1 = torch.cat([e[1] for e in ctx.undo.entries])
            o2 = torch.cat([e[2] for e in ctx.undo.entries])
            uniq, inv = torch.unique(ys, return_inverse=True)
            first = torch.full((uniq.numel(),), ys.numel(), dtype=torch.long, device=ys.device)
            first.scatter_reduce_(0, inv, torch.arange(ys.numel(), device=ys.device), "amin")
            flat = idx.reshape(-1).long()
            pos = torch.searchsorted(uniq, flat).clamp_(max=uniq.numel() - 1)
            hit = (uniq[pos] == flat).nonzero().flatten()
            if hit.numel():
                rows = flat[hit]
                src = first[pos[hit]]
                bb = hit // cols
                g1.index_add_(0, bb, coef1.reshape(-1)[hit].unsqueeze(1) * (o2[src] - mem.memory_v2[rows]))
                g2.index_add_(0, bb, coef2.reshape(-1)[hit].unsqueeze(1) * (o1[src] - mem.memory_v1[rows]))
        return g1, g2, None, None, None


class _FusedLossFn(torch.autograd.Function):
    """CRDLoss hot path: loss and dL/dv1, dL/dv2 from one pass over the rows."""

    @staticmethod
    def forward(ctx, v1, v2, mem, idx, n_data):
        loss, _, g1, g2, _, _ = crd_fused_loss_grad(
            mem.memory_v1, mem.memory_v2, v1, v2, idx, mem._T, mem.params[2:4], n_data, mem._K)
        ctx.save_for_backward(g1, g2)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loss):
        g1, g2 = ctx.saved_tensors       # repeatable: nothing here touches the (already updated) banks
        return grad_loss * g1, grad_loss * g2, None, None, None


class ContrastMemory(nn.Module):
    """memory buffer that supplies large amount of negative samples."""

    def __init__(self, inputSize, outputSize, K, T=0.07, momentum=0.5, device=None):
        """`device` (extension, default None = the reference's behaviour: buffers drawn on the host from the global torch
        RNG in the reference's order) draws the banks directly on that device -- for banks too large to stage through host
        memory (16M x 128 x 2 = 16 GB at BASELINE config 5)."""
        super(ContrastMemory, self).__init__()
        self.nLem = outputSize
        self.unigrams = torch.ones(self.nLem)
        self.multinomial = AliasMethod(self.unigrams)
        self.K = K
        self.register_buffer('params', torch.tensor([K, T, -1, -1, momentum], device=device))
        stdv = 1. / math.sqrt(inputSize / 3)
        self.register_buffer('memory_v1', torch.rand(outputSize, inputSize, device=device).mul_(2 * stdv).add_(-stdv))
        self.register_buffer('memory_v2', torch.rand(outputSize, inputSize, device=device).mul_(2 * stdv).add_(-stdv))
        if device is not None and torch.device(device).type == "cuda":
            self.multinomial.cuda(device)
        self._refresh_scalars()
        self._pending = weakref.WeakSet()     # undo logs of differentiable forwards still alive

    # -- host-side scalar cache (replaces the five .item() syncs of :26-31) --
    def _refresh_scalars(self):
        p = self.params.detach().cpu()
        self._K = int(p[0].item())
        self._T = p[1].item()               # fp32(0.07) widened, exactly what :27 reads
        self._momentum = p[4].item()
        self._z_ready = bool(p[2].item() > 0 and p[3].item() > 0)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._refresh_scalars()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if self.params.is_cuda and not self.multinomial.prob.is_cuda:
            self.multinomial.cuda(self.params.device)     # the reference calls .cuda() in __init__ (:17)
        return out

    # -- pieces shared with CRDLoss's fused path --
    def _check_inputs(self, v1, v2, y, idx):
        if not (v1.is_cuda and self.memory_v1.is_cuda):
            raise RuntimeError("ContrastMemory runs on CUDA tensors only (move the module with .cuda()/.to(device))")
        B = v1.size(0)
        v1, v2, y = _as_f32(v1), _as_f32(v2), _as_i64(y)
        if idx is None:                                              # :37-39
            idx = self.multinomial.draw(B * (self.K + 1), y=y, cols=self.K + 1).view(B, -1)
        elif idx.dtype == torch.int32:      # extension: int32 row ids (n_data < 2^31) halve the index traffic
            idx = idx.contiguous()
        else:
            idx = _as_i64(idx)
        idx = idx.view(B, self._K + 1)          # same RuntimeError as :42 when idx has the wrong width
        return v1, v2, y, idx

    def _ensure_Z(self, v1, v2, idx):
        """First-call Monte-Carlo normaliser (:52-59): Z = mean(exp(dot/T)) * outputSize."""
        if self._z_ready:
            return
        crd_scores(self.memory_v1, self.memory_v2, v1.detach(), v2.detach(), idx, self._T,
                   set_Z=self.params[2:4], want_out=False)
        z1, z2 = self.params[2].item(), self.params[3].item()
        print("normalization constant Z_v1 is set to {:.1f}".format(z1))
        print("normalization constant Z_v2 is set to {:.1f}".format(z2))
        self._z_ready = True

    def _update(self, v1, v2, y):
        """:66-79, applied after every read of this step."""
        with torch.no_grad():
            if len(self._pending):
                old1, old2 = self.memory_v1[y], self.memory_v2[y]
                for log in list(self._pending):
                    log.entries.append((y, old1, old2))
            crd_memory_update(self.memory_v1, self.memory_v2, v1.detach(), v2.detach(), y, self._momentum)

    def forward(self, v1, v2, y, idx=None):
        v1, v2, y, idx = self._check_inputs(v1, v2, y, idx)
        self._ensure_Z(v1, v2, idx)
        if torch.is_grad_enabled() and (v1.requires_grad or v2.requires_grad):
            undo = _UndoLog()
            self._pending.add(undo)
            out_v1, out_v2 = _ScoresFn.apply(v1, v2, self, idx, undo)
        else:
            o1, o2, _ = crd_scores(self.memory_v1, self.memory_v2, v1, v2, idx, self._T, Z=self.params[2:4])
            out_v1, out_v2 = o1.unsqueeze(2), o2.unsqueeze(2)
        self._update(v1, v2, y)
        return out_v1, out_v2

    def fused_nce_loss(self, v1, v2, y, idx, n_data):
        """criterion_s(out_v1) + criterion_t(out_v2) of CRDLoss.forward (:184-187)
        without materialising out_v1/out_v2; same side effects as `forward`."""
        v1, v2, y, idx = self._check_inputs(v1, v2, y, idx)
        self._ensure_Z(v1, v2, idx)
        loss = _FusedLossFn.apply(v1, v2, self, idx, n_data)
        self._update(v1, v2, y)
        return loss


# --------------------------------------------------------------------------- #
# CRDLoss / ContrastLoss / Embed / Normalize
# --------------------------------------------------------------------------- #
_HEAD_STREAMS = {}


def embed_pair(embed_s, f_s, embed_t, f_t):
    """The two Embed heads are independent chains of small GEMMs (each leaves most SMs idle): the teacher head runs on a
    second stream beside the student head.  Autograd replays every backward node on the stream of its forward, so the two
    backward chains (the longer half: dX, dW, bias reductions per layer) overlap as well -- forward and backward of the heads
    cost about half the time, with bit-identical results.  Works inside a CUDA-graph capture (fork / join become graph
    dependencies).  MML_HEADS_STREAMS=0 keeps them in line."""
    if not (f_s.is_cuda and f_t.is_cuda) or os.environ.get("MML_HEADS_STREAMS", "1") != "1":
        return embed_s(f_s), embed_t(f_t)
    dev = f_t.device
    side = _HEAD_STREAMS.get(dev)
    if side is None:
        side = _HEAD_STREAMS[dev] = torch.cuda.Stream(dev, priority=-1)
    cur = torch.cuda.current_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        v_t = embed_t(f_t)
    v_s = embed_s(f_s)
    cur.wait_stream(side)
    v_t.record_stream(cur)
    return v_s, v_t


class CRDLoss(nn.Module):
    """CRD Loss function
    includes two symmetric parts:
    (a) using teacher as anchor, choose positive and negatives over the student side
    (b) using student as anchor, choose positive and negatives over the teacher side

    Args (CRD_criterion.py:150-157):
        opt.s_dim / opt.t_dim: the dimension of student's / teacher's feature
        opt.feat_dim: the dimension of the projection space
        opt.nce_k / opt.nce_t / opt.nce_m: negatives per positive, temperature, memory momentum
        opt.n_data: number of training samples = rows of each memory bank
    """

    def __init__(self, opt):
        super(CRDLoss, self).__init__()
        self.embed_s = Embed(opt.s_dim, opt.feat_dim)
        self.embed_t = Embed(opt.t_dim, opt.feat_dim)
        self.contrast = ContrastMemory(opt.feat_dim, opt.n_data, opt.nce_k, opt.nce_t, opt.nce_m,
                                       device=getattr(opt, "bank_device", None))     # optional, see ContrastMemory
        self.criterion_t = ContrastLoss(opt.n_data)
        self.criterion_s = ContrastLoss(opt.n_data)

    def forward(self, f_s, f_t, idx, contrast_idx=None):
        """
        f_s: [batch_size, s_dim] student feature;  f_t: [batch_size, t_dim] teacher feature
        idx: [batch_size] dataset indices of the positives
        contrast_idx: [batch_size, nce_k + 1] indices (column 0 = positive), or None to sample
        Returns the contrastive loss, shape [1].
        """
        f_s, f_t = embed_pair(self.embed_s, f_s, self.embed_t, f_t)
        if self.criterion_s.n_data != self.criterion_t.n_data:
            out_s, out_t = self.contrast(f_s, f_t, idx, contrast_idx)
            return self.criterion_s(out_s) + self.criterion_t(out_t)
        return self.contrast.fused_nce_loss(f_s, f_t, idx, contrast_idx, self.criterion_s.n_data)


class ContrastLoss(nn.Module):
    """contrastive loss, corresponding to Eq (18) -- stand-alone form for callers that
    hold out_v1/out_v2 (CRDLoss uses the fused kernel instead)."""

    def __init__(self, n_data):
        super(ContrastLoss, self).__init__()
        self.n_data = n_data

    def forward(self, x):
        bsz = x.shape[0]
        m = x.size(1) - 1
        Pn = 1 / float(self.n_data)
        noise = m * Pn
        pos = x.select(1, 0)
        neg = x.narrow(1, 1, m)
        log_D1 = (pos / (pos + (noise + eps))).log()
        log_D0 = (noise / (neg + (noise + eps))).log()
        return -(log_D1.sum(0) + log_D0.reshape(-1, 1).sum(0)) / bsz


class Embed(nn.Module):
    """Embedding module: Linear-ReLU-Linear then L2 normalisation (:219-233).
    `layers=1` gives the single-Linear head of `MIA 2022/CL_utils/CRD_criterion.py:223`."""

    def __init__(self, dim_in=1024, dim_out=128, layers=2):
        super(Embed, self).__init__()
        if layers == 1:
            self.linear = nn.Linear(dim_in, dim_out)
        else:
            self.linear = nn.Sequential(nn.Linear(dim_in, dim_out), nn.ReLU(), nn.Linear(dim_out, dim_out))
        self.l2norm = Normalize(2)

    def forward(self, x):
        x = x.view(x.shape[0], -1)
        x = self.linear(x)
        return self.l2norm(x)


class _L2NormFn(torch.autograd.Function):
    """Row-wise x / ||x||_2 as one kernel forward and one backward (`mml_l2norm_fwd/bwd`)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        B, D = x.shape
        y = torch.empty_like(x)
        nrm = torch.empty(B, dtype=torch.float32, device=x.device)
        _cabi.check(_cabi.lib().mml_l2norm_fwd(_cabi.dptr(x), B, D, _cabi.dptr(y), _cabi.dptr(nrm), _cabi.cur_stream(x.device)),
                    "mml_l2norm_fwd")
        ctx.save_for_backward(y, nrm)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        y, nrm = ctx.saved_tensors
        gy = gy.contiguous()
        gx = torch.empty_like(gy)
        _cabi.check(_cabi.lib().mml_l2norm_bwd(_cabi.dptr(gy), _cabi.dptr(y), _cabi.dptr(nrm), y.shape[0], y.shape[1],
                                               _cabi.dptr(gx), _cabi.cur_stream(gy.device)), "mml_l2norm_bwd")
        return gx


_ALLOW_HOST_NORMALIZE = False      # set by tests/test_sharded_cpu.py only: the gloo choreography test runs the heads on the host


class Normalize(nn.Module):
    """normalization layer (no epsilon, as the reference).  power=2 on a CUDA fp32 matrix runs as one fused kernel
    (forward) / one (backward); other powers / ranks keep the reference's op chain on the GPU.  CUDA tensors only, like
    every module of this package (the gloo CPU tests of the sharded choreography substitute their own backend)."""

    def __init__(self, power=2):
        super(Normalize, self).__init__()
        self.power = power

    def forward(self, x):
        if self.power == 2 and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2:
            return _L2NormFn.apply(x)
        if not x.is_cuda and not _ALLOW_HOST_NORMALIZE:
            raise RuntimeError("Normalize runs on CUDA tensors only (no CPU fallback)")
        norm = x.pow(self.power).sum(1, keepdim=True).pow(1. / self.power)
        return x.div(norm)
