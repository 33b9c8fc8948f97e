"""CPU oracle for the gated Kronecker fusion path.  TEST INFRASTRUCTURE ONLY.

Functional restatement (torch-CPU fp32) of the reference's
`MICCAI-2022/fusion.py` (`BilinearFusion` :6-63, `TrilinearFusion_A` :66-132,
`TrilinearFusion_B` :135-201), `utils.py:239-244` (`init_max_weights`) and
`KD_loss.py:7-17` (`DistillKL`).  The Kronecker tensor IS materialised here, on
purpose: this is the checker for the never-materialise CUDA kernels.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import it.

Parity status: PINNED against the unmodified reference modules through the
fixtures `oracle/make_golden.py` writes under `tests/golden/` (eval mode, and
train mode with dropout p=0 so BatchNorm batch statistics are covered; torch's
dropout mask stream is device/version specific, so masks themselves are
"parity unpinned" -- see DESIGN.md).

State dicts use the reference's key names (`linear_h1.0.weight`,
`linear_z1.weight`, `encoder1.0.weight`, `encoder1.1.running_mean`, ...).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def init_max_weights_(sd: dict, generator: torch.Generator | None = None) -> None:
    """utils.py:239-244 applied to a state dict: every nn.Linear weight ~
    N(0, 1/sqrt(fan_in)), bias 0.  Linear layers are the 2-D `*.weight` entries
    (nn.Bilinear weights are 3-D, BatchNorm weights 1-D: both keep torch defaults)."""
    for k in list(sd.keys()):
        if k.endswith("weight") and sd[k].dim() == 2:
            stdv = 1.0 / math.sqrt(sd[k].size(1))
            sd[k].normal_(0, stdv, generator=generator)
            sd[k[:-6] + "bias"].zero_()


class ExactOps:
    """The two contractions the tensor-core kernels replace, as the reference computes them.  Tests pass a variant that
    rounds the operands to TF32 first (forward and backward) to get the error ANY TF32 implementation of these
    contractions carries on a given fixture -- the yardstick for tolerances above the nominal 2e-3."""

    @staticmethod
    def linear(A, W, b):
        return F.linear(A, W, b)

    @staticmethod
    def bilinear(a, b, W, bias):
        return F.bilinear(a, b, W, bias)


def _gate(sd, tag, own, a, b, gated, use_bilinear, ops=ExactOps):
    """One gated multimodal unit, fusion.py:41-46: h = ReLU(Linear(own));
    z = Bilinear(a, b) or Linear(cat(a, b)); o = ReLU(Linear(sigmoid(z) * h)).
    (linear_o's trailing Dropout is identity in eval / p=0.)"""
    if gated:
        h = torch.relu(F.linear(own, sd[f"linear_h{tag}.0.weight"], sd[f"linear_h{tag}.0.bias"]))
        if use_bilinear:
            z = ops.bilinear(a, b, sd[f"linear_z{tag}.weight"], sd[f"linear_z{tag}.bias"])
        else:
            z = F.linear(torch.cat((a, b), dim=1), sd[f"linear_z{tag}.0.weight"], sd[f"linear_z{tag}.0.bias"])
        pre = torch.sigmoid(z) * h
    else:
        pre = own
    return torch.relu(F.linear(pre, sd[f"linear_o{tag}.0.weight"], sd[f"linear_o{tag}.0.bias"]))


def _append_one(o):
    """fusion.py:56-57 -- constant 1 in the LAST slot of each factor."""
    return torch.cat((o, torch.ones(o.shape[0], 1, dtype=o.dtype, device=o.device)), 1)


def kron_rows(*factors):
    """fusion.py:58 / :126-127 -- batched outer product flattened row-major over
    (i, j[, l]): out[b, (i*(d2+1)+j)*(d3+1)+l]."""
    out = factors[0]
    for f in factors[1:]:
        out = torch.bmm(out.unsqueeze(2), f.unsqueeze(1)).flatten(start_dim=1)
    return out


def _bn(sd, name, x, training):
    """nn.BatchNorm1d defaults (momentum 0.1, eps 1e-5); mutates running stats
    in `sd` when training, as the module does."""
    if training:
        sd[f"{name}.num_batches_tracked"] += 1
    return F.batch_norm(x, sd[f"{name}.running_mean"], sd[f"{name}.running_var"],
                        sd[f"{name}.weight"], sd[f"{name}.bias"], training, 0.1, 1e-5)


def bilinear_fusion_forward(sd, vec1, vec2, *, skip=1, use_bilinear=1, gate1=1, gate2=1,
                            training=False, ops=ExactOps):
    """BilinearFusion.forward, fusion.py:36-63, with every Dropout as identity
    (eval mode, or train mode with dropout_rate=0)."""
    vec1 = torch.relu(vec1)                 # :38
    vec2 = torch.relu(vec2)                 # :39
    o1 = _gate(sd, 1, vec1, vec1, vec2, gate1, use_bilinear, ops)
    o2 = _gate(sd, 2, vec2, vec1, vec2, gate2, use_bilinear, ops)
    o1, o2 = _append_one(o1), _append_one(o2)
    o12 = kron_rows(o1, o2)                 # :58
    out = ops.linear(o12, sd["encoder1.0.weight"], sd["encoder1.0.bias"])
    out = torch.relu(_bn(sd, "encoder1.1", out, training))
    if skip:
        out = torch.cat((out, o1, o2), 1)   # :61
    out = F.linear(out, sd["encoder2.0.weight"], sd["encoder2.0.bias"])
    return torch.relu(_bn(sd, "encoder2.1", out, training))


def polynomial_fusion_forward(sd, vec1, vec2, *, skip=1, use_bilinear=1, gate1=1, gate2=1, training=False, ops=ExactOps):
    """PolynomialFusion.forward, `MIA 2023/stage2_unimodal_student/fusion.py:38-72` (every Dropout as identity):
    BilinearFusion's gated Kronecker + encoder1, then a SECOND Kronecker of [encoder1_out, 1] with itself (:64-68)."""
    vec1 = torch.relu(vec1)                 # :40
    vec2 = torch.relu(vec2)                 # :41
    o1 = _gate(sd, 1, vec1, vec1, vec2, gate1, use_bilinear, ops)
    o2 = _gate(sd, 2, vec2, vec1, vec2, gate2, use_bilinear, ops)
    o1, o2 = _append_one(o1), _append_one(o2)
    o12 = kron_rows(o1, o2)                 # :60
    out12 = ops.linear(o12, sd["encoder1.0.weight"], sd["encoder1.0.bias"])
    out12 = _append_one(torch.relu(_bn(sd, "encoder1.1", out12, training)))     # :62-64
    o1212 = kron_rows(out12, out12)         # :65
    out = ops.linear(o1212, sd["encoder2.0.weight"], sd["encoder2.0.bias"])
    out = torch.relu(_bn(sd, "encoder2.1", out, training))
    if skip:
        out = torch.cat((out, o1, o2), 1)   # :68
    out = F.linear(out, sd["encoder3.0.weight"], sd["encoder3.0.bias"])
    return torch.relu(_bn(sd, "encoder3.1", out, training))


def trilinear_fusion_forward(sd, vec1, vec2, vec3, *, variant="A", skip=1, use_bilinear=1,
                             gate1=1, gate2=1, gate3=1):
    """TrilinearFusion_A.forward (fusion.py:99-132) / _B (:168-201), Dropout as
    identity.  No input ReLU, no BatchNorm.  Variant B gates the graph branch
    with (vec2, vec1) instead of (vec2, vec3) (:179 vs :110)."""
    o1 = _gate(sd, 1, vec1, vec1, vec3, gate1, use_bilinear)
    if variant == "A":
        o2 = _gate(sd, 2, vec2, vec2, vec3, gate2, use_bilinear)
    else:
        o2 = _gate(sd, 2, vec2, vec2, vec1, gate2, use_bilinear)
    o3 = _gate(sd, 3, vec3, vec1, vec3, gate3, use_bilinear)
    o1, o2, o3 = _append_one(o1), _append_one(o2), _append_one(o3)
    o123 = kron_rows(o1, o2, o3)            # :126-127
    out = torch.relu(F.linear(o123, sd["encoder1.0.weight"], sd["encoder1.0.bias"]))
    if skip:
        out = torch.cat((out, o1, o2, o3), 1)
    return torch.relu(F.linear(out, sd["encoder2.0.weight"], sd["encoder2.0.bias"]))


def kron_linear(factors, weight, bias):
    """The contraction the CUDA kernel replaces: (append-1 Kronecker rows) @ W^T + b,
    computed in float64 from fp32 inputs (factors WITHOUT the appended 1)."""
    aug = [_append_one(f).double() for f in factors]
    return kron_rows(*aug) @ weight.double().t() + bias.double()


def distill_kl(y_s, y_t, T):
    """KD_loss.py:13-17 -- KL(softmax(y_t/T) || softmax(y_s/T)) * T^2 / B, sum reduction."""
    p_s = F.log_softmax(y_s / T, dim=1)
    p_t = F.softmax(y_t / T, dim=1)
    return F.kl_div(p_s, p_t, reduction="sum") * (T ** 2) / y_s.shape[0]


def kron_dropout_mask(seed: int, B: int, Kk: int, p: float) -> torch.Tensor:
    """The counter-based `post_fusion_dropout` multiplier this repo DEFINES for the never-stored
    Kronecker tensor (csrc/kron_common.cuh; torch's Philox mask stream cannot apply to a tensor
    that does not exist, so mask parity with the reference is "unpinned" by construction).
    Returns float32 [B, Kk] with entries 0 or 65536/(65536 - round(p*65536))."""
    import numpy as np
    return kron_dropout_mask_at(seed, np.arange(B)[:, None], np.arange(Kk)[None, :], Kk, p)


def _kron_hash(c, seed: int):
    """csrc/kron_common.cuh kron_hash on a uint64 counter array -> uint32."""
    import numpy as np
    lo = (c & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (c >> np.uint64(32)).astype(np.uint32)
    s_lo, s_hi = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        h = lo ^ s_lo
        h = h * np.uint32(0x9E3779B1)
        h = h ^ (hi ^ s_hi)
        h = h ^ (h >> np.uint32(16))
        h = h * np.uint32(0x7FEB352D)
        h = h ^ (h >> np.uint32(15))
        h = h * np.uint32(0x846CA68B)
        h = h ^ (h >> np.uint32(16))
    return h


def kron_dropout_mask_at(seed: int, rows, cols, Kk: int, p: float) -> torch.Tensor:
    """Same mask evaluated only at (rows, cols) -- numpy-broadcastable integer arrays of batch rows b and flattened
    Kronecker columns k -- so that tests at BASELINE sizes need not build the whole [B, Kk] mask.

    Definition (kron_common.cuh): element (b, k) is DROPPED iff bit (b & 31) of word (b >> 5, k) is set; a word is the
    bit-sliced comparison U < thresh of 32 independent 16-bit uniforms, thresh = round(p * 65536): from the lowest set bit
    i of thresh upwards, lt = (plane_i | lt) if bit i of thresh is set else (plane_i & lt), with
    plane_i = hash((((b >> 5) * Kk + k) << 4) + i, seed)."""
    import numpy as np
    rows = np.asarray(rows, dtype=np.uint64)
    cols = np.asarray(cols, dtype=np.uint64)
    shape = np.broadcast(rows, cols).shape
    thresh = min(int(p * 65536.0 + 0.5), 65535) if p > 0 else 0
    if thresh == 0:
        return torch.ones(shape)
    c0 = ((rows >> np.uint64(5)) * np.uint64(Kk) + cols) << np.uint64(4)
    lt = np.zeros(shape, dtype=np.uint32)
    started = False
    for i in range(16):
        bit = (thresh >> i) & 1
        if not started and not bit:
            continue                     # planes below the lowest set bit cannot change the outcome
        started = True
        plane = _kron_hash(np.broadcast_to(c0 + np.uint64(i), shape), seed)
        lt = (plane | lt) if bit else (plane & lt)
    dropped = (lt >> (np.broadcast_to(rows, shape) & np.uint64(31)).astype(np.uint32)) & np.uint32(1)
    scale = np.float32(65536.0) / np.float32(65536 - thresh)
    return torch.from_numpy(np.where(dropped == 0, scale, np.float32(0)).astype(np.float32))
