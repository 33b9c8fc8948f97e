"""Gated Kronecker fusion -- host-side mirror of the reference's `MICCAI-2022/fusion.py`
(`BilinearFusion` :6-63, `TrilinearFusion_A` :66-132, `TrilinearFusion_B` :135-201) with the
same constructor keywords, `forward` signatures, sub-module / state_dict names and RNG
consumption order (same seed => same initial weights).

What runs where:
  * gates (`linear_h*`, `linear_z*`, `linear_o*`), BatchNorm/ReLU/Dropout on [B, mmhid] and
    `encoder2` are small and stay host-side PyTorch;
  * the hot contraction -- append-1, outer product(s), flatten, `post_fusion_dropout`,
    `encoder1[0]` -- is ONE kernel family that never materialises the (d+1)^2 / (d+1)^3 tensor:
    forward, weight gradient and factor gradients all on tcgen05 tensor cores (`mml_kron_linear_fwd`,
    `mml_kron_linear_wgrad`, `mml_kron_linear_dgrad`); exact-fp32 CUDA-core kernels (`*_simt`) for shapes the
    tensor kernels do not take and as the on-GPU cross-check of the tests.
  * `post_fusion_dropout` masks come from a counter-based hash of (seed, b, k) instead of
    torch's Philox stream (a tensor that is never stored cannot carry a torch mask); eval mode
    and p=0 are bit-for-bit the reference's math.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import _cabi


def ctypes_ptr(addr):
    import ctypes
    return ctypes.c_void_p(addr)


def init_max_weights(module):
    """utils.py:239-244 of the reference: every nn.Linear gets N(0, 1/sqrt(fan_in)) weights, zero bias."""
    for m in module.modules():
        if type(m) == nn.Linear:
            stdv = 1. / math.sqrt(m.weight.size(1))
            m.weight.data.normal_(0, stdv)
            m.bias.data.zero_()
        if hasattr(m, "invalidate_kron_caches"):      # `.data` writes do not bump the version the packed copies are keyed on
            m.invalidate_kron_caches()


# --------------------------------------------------------------------------- #
# the Kronecker linear op
# --------------------------------------------------------------------------- #
_BRANCH_STREAMS = {}       # device -> side streams of the gated branches


class KronLinearState:
    """Per-module cache: chunk table (device) and the packed TF32 copies of the weight, refreshed whenever the dense
    weight's version counter or storage changes.  Writes through `.data` (`init_max_weights`, `p.data.clamp_()`, EMA
    updates) do not bump the version counter: call `invalidate()` after them (the modules' `init_max_weights`,
    `load_state_dict` and `_apply` hooks do)."""

    def __init__(self, dims):
        self.dims = tuple(int(d) for d in dims) + ((0,) if len(dims) == 2 else ())
        self.table = None
        self.packed = None
        self.packed_key = None
        self.path = "auto"          # "auto" | "simt" (tests use "simt" as the exact-fp32 cross-check)
        self._plans = {}            # (B, N) -> (tensor-core path usable, workspace bytes)
        self._wg_plans = {}         # (B, N) -> (tensor-core wgrad usable, workspace bytes)
        self._dg_plans = {}         # (B, N) -> (tensor-core dgrad usable, workspace bytes)
        self.packed_t = None        # transposed packed weight (dgrad's B operand)
        self.last_stats = None      # per-tile column sums of the last forward (consumed by the BatchNorm finisher)
        self.packed_t_key = None

    def invalidate(self):
        """Forget the packed weight copies (next forward / backward repacks from the dense weight)."""
        self.packed_key = None
        self.packed_t_key = None

    def dgrad_plan(self, B, N):
        key = (B, N)
        if key not in self._dg_plans:
            lib = _cabi.lib()
            ok = bool(lib.mml_kron_dgrad_supported(B, N, *self.dims))
            self._dg_plans[key] = (ok, lib.mml_kron_dgrad_workspace_bytes(B, N, *self.dims) if ok else 0)
        return self._dg_plans[key]

    def ensure_t(self, weight, key=None):
        lib = _cabi.lib()
        d1, d2, d3 = self.dims
        dev = weight.device
        self.ensure_table(dev)
        key = key if key is not None else (weight.data_ptr(), weight._version, tuple(weight.shape))
        if key != self.packed_t_key:
            N = weight.shape[0]
            nfl = lib.mml_kron_packed_t_floats(N, d1, d2, d3)
            if self.packed_t is None or self.packed_t.numel() != nfl or self.packed_t.device != dev:
                self.packed_t = torch.empty(nfl, dtype=torch.float32, device=dev)
            _cabi.check(lib.mml_kron_pack_weight_t(_cabi.dptr(weight.detach()), N, d1, d2, d3, _cabi.dptr(self.table),
                                                   _cabi.dptr(self.packed_t), _cabi.cur_stream(dev)), "mml_kron_pack_weight_t")
            self.packed_t_key = key

    def wgrad_plan(self, B, N):
        key = (B, N)
        if key not in self._wg_plans:
            lib = _cabi.lib()
            ok = bool(lib.mml_kron_wgrad_supported(B, N, *self.dims))
            self._wg_plans[key] = (ok, lib.mml_kron_wgrad_workspace_bytes(B, N, *self.dims) if ok else 0)
        return self._wg_plans[key]

    def ensure_table(self, device):
        if self.table is None or self.table.device != device:
            lib = _cabi.lib()
            d1, d2, d3 = self.dims
            n = lib.mml_kron_num_chunks(d1, d2, d3)
            host = torch.empty(n * 8, dtype=torch.int32)
            _cabi.check(lib.mml_kron_chunk_table_host(d1, d2, d3, _cabi.hptr(host)), "mml_kron_chunk_table_host")
            self.table = host.to(device)
            self.packed_key = None
            self.packed_t_key = None

    def plan(self, B, N):
        key = (B, N)
        if key not in self._plans:
            lib = _cabi.lib()
            ok = bool(lib.mml_kron_fwd_supported(B, N, *self.dims))
            self._plans[key] = (ok, lib.mml_kron_fwd_workspace_bytes(B, N, *self.dims) if ok else 0)
        return self._plans[key]

    def ensure(self, weight, key=None):
        lib = _cabi.lib()
        d1, d2, d3 = self.dims
        dev = weight.device
        self.ensure_table(dev)
        key = key if key is not None else (weight.data_ptr(), weight._version, tuple(weight.shape))
        if key != self.packed_key:
            N = weight.shape[0]
            nfl = lib.mml_kron_packed_floats(N, d1, d2, d3)
            if self.packed is None or self.packed.numel() != nfl or self.packed.device != dev:
                self.packed = torch.empty(nfl, dtype=torch.float32, device=dev)
            _cabi.check(lib.mml_kron_pack_weight(_cabi.dptr(weight.detach()), N, d1, d2, d3, _cabi.dptr(self.table),
                                                 _cabi.dptr(self.packed), _cabi.cur_stream(dev)), "mml_kron_pack_weight")
            self.packed_key = key


class _KronLinearFn(torch.autograd.Function):
    """y = kron(append1(f1), append1(f2)[, append1(f3)]) * mask @ W^T + bias, without the Kronecker tensor."""

    @staticmethod
    def forward(ctx, state, weight, bias, drop_p, training, seed, wkey, want_stats, *factors):
        lib = _cabi.lib()
        state.last_stats = None
        d1, d2, d3 = state.dims
        fs = [f.contiguous() for f in factors]
        B = fs[0].shape[0]
        N = weight.shape[0]
        dev = fs[0].device
        st = _cabi.cur_stream(dev)
        y = torch.empty(B, N, dtype=torch.float32, device=dev)
        f3p = _cabi.dptr(fs[2]) if d3 > 0 else None
        w = weight.detach().contiguous()
        bptr = _cabi.dptr(bias.detach().contiguous()) if bias is not None else None
        seed_t = seed if torch.is_tensor(seed) else None        # device word: fresh mask on every CUDA-graph replay
        seed = 0 if seed_t is not None else int(seed)
        sdev = _cabi.dptr(seed_t, torch.int64) if seed_t is not None else None
        tc_ok, nws = state.plan(B, N)
        if state.path == "auto" and tc_ok:
            state.ensure(weight, wkey)
            ws = torch.empty(nws, dtype=torch.uint8, device=dev)
            tiles = int(lib.mml_kron_fwd_stat_tiles(B, N, d1, d2, d3)) if want_stats else 0
            stats = torch.empty(tiles, 2, N, dtype=torch.float32, device=dev) if tiles > 0 else None
            rc = lib.mml_kron_linear_fwd_stats(_cabi.dptr(fs[0]), _cabi.dptr(fs[1]), f3p, B, d1, d2, d3, _cabi.dptr(state.table),
                                               _cabi.dptr(state.packed), bptr, N, float(drop_p), seed, sdev, int(training),
                                               _cabi.dptr(y), _cabi.dptr(stats), _cabi.dptr(ws), nws, st)
            _cabi.check(rc, "mml_kron_linear_fwd_stats")
            state.last_stats = stats         # per-tile column sums of y, y^2 from the epilogue (BatchNorm finisher)
        else:
            rc = lib.mml_kron_linear_fwd_simt(_cabi.dptr(fs[0]), _cabi.dptr(fs[1]), f3p, B, d1, d2, d3, _cabi.dptr(w), bptr,
                                              N, float(drop_p), seed, sdev, int(training), _cabi.dptr(y), st)
            _cabi.check(rc, "mml_kron_linear_fwd_simt")
        ctx.save_for_backward(w, *fs)
        ctx.state = state
        ctx.wkey = wkey
        ctx.cfg = (state.dims, float(drop_p), int(training), seed, bias is not None)
        ctx.seed_t = seed_t
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = _cabi.lib()
        w, *fs = ctx.saved_tensors
        (d1, d2, d3), drop_p, training, seed, has_bias = ctx.cfg
        sdev = _cabi.dptr(ctx.seed_t, torch.int64) if ctx.seed_t is not None else None
        dy = dy.contiguous()
        B, N = dy.shape
        dev = dy.device
        need_w = ctx.needs_input_grad[1]
        need_f = any(ctx.needs_input_grad[8:])
        dW = torch.empty_like(w) if need_w else None
        dfs = [torch.empty_like(f) for f in fs] if need_f else [None] * len(fs)
        state = ctx.state
        f3p = _cabi.dptr(fs[2]) if d3 > 0 else None
        st = _cabi.cur_stream(dev)
        wg_ok, wg_ws = state.wgrad_plan(B, N)
        simt_dW = dW
        if need_w and state.path == "auto" and wg_ok:            # weight gradient on the tensor cores
            state.ensure_table(dev)
            ws = torch.empty(wg_ws + 1024, dtype=torch.uint8, device=dev)
            off = (-ws.data_ptr()) % 1024
            rc = lib.mml_kron_linear_wgrad(_cabi.dptr(fs[0]), _cabi.dptr(fs[1]), f3p, B, d1, d2, d3, _cabi.dptr(state.table),
                                           _cabi.dptr(dy), N, drop_p, seed, sdev, training, _cabi.dptr(dW),
                                           ctypes_ptr(ws.data_ptr() + off), wg_ws, st)
            _cabi.check(rc, "mml_kron_linear_wgrad")
            simt_dW = None
        dg_ok, dg_ws = state.dgrad_plan(B, N)
        if need_f and state.path == "auto" and dg_ok:            # factor gradients on the tensor cores
            state.ensure_t(w, ctx.wkey)
            ws = torch.empty(dg_ws, dtype=torch.uint8, device=dev)
            rc = lib.mml_kron_linear_dgrad(_cabi.dptr(fs[0]), _cabi.dptr(fs[1]), f3p, B, d1, d2, d3, _cabi.dptr(state.table),
                                           _cabi.dptr(state.packed_t), _cabi.dptr(dy), N, drop_p, seed, sdev, training,
                                           _cabi.dptr(dfs[0]), _cabi.dptr(dfs[1]), _cabi.dptr(dfs[2]) if d3 > 0 else None,
                                           _cabi.dptr(ws), dg_ws, st)
            _cabi.check(rc, "mml_kron_linear_dgrad")
            need_f_simt = False
        else:
            need_f_simt = need_f
        if need_f_simt or simt_dW is not None:
            sd = dfs if need_f_simt else [None] * len(fs)
            rc = lib.mml_kron_linear_bwd_simt(
                _cabi.dptr(fs[0]), _cabi.dptr(fs[1]), f3p, B, d1, d2, d3, _cabi.dptr(w), _cabi.dptr(dy), N, drop_p, seed,
                sdev, training, _cabi.dptr(sd[0]), _cabi.dptr(sd[1]), _cabi.dptr(sd[2]) if d3 > 0 else None,
                _cabi.dptr(simt_dW), st)
            _cabi.check(rc, "mml_kron_linear_bwd_simt")
        dbias = dy.sum(0) if (has_bias and ctx.needs_input_grad[2]) else None
        return (None, dW, dbias, None, None, None, None, None, *dfs)


class _BNReLUFn(torch.autograd.Function):
    """relu(batch_norm(y)) in training mode through `mml_bn_relu_fwd`: batch statistics from the Kronecker kernel's epilogue
    partials when it emitted them (else from y), running-statistics update, normalise + affine + ReLU in one pass.
    Backward: the ReLU mask and torch's native batch-norm backward on [B, N] (small; SURVEY.md §7.2 keeps it host-side)."""

    @staticmethod
    def forward(ctx, y, stats, weight, bias, bn):
        lib = _cabi.lib()
        y = y.contiguous()
        B, N = y.shape
        out = torch.empty_like(y)
        mean = torch.empty(N, dtype=torch.float32, device=y.device)
        invstd = torch.empty(N, dtype=torch.float32, device=y.device)
        track = bn.track_running_stats and bn.running_mean is not None
        rc = lib.mml_bn_relu_fwd(_cabi.dptr(y), B, N, _cabi.dptr(stats), 0 if stats is None else stats.shape[0],
                                 _cabi.dptr(weight.detach()), _cabi.dptr(bias.detach()),
                                 _cabi.dptr(bn.running_mean) if track else None, _cabi.dptr(bn.running_var) if track else None,
                                 float(bn.momentum), float(bn.eps), _cabi.dptr(out), _cabi.dptr(mean), _cabi.dptr(invstd),
                                 _cabi.cur_stream(y.device))
        _cabi.check(rc, "mml_bn_relu_fwd")
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        ctx.save_for_backward(y, out, mean, invstd, weight)
        ctx.eps = float(bn.eps)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        y, out, mean, invstd, weight = ctx.saved_tensors
        g = (g * (out > 0)).contiguous()
        dy, dw, db = torch.ops.aten.native_batch_norm_backward(g, y, weight, None, None, mean, invstd, True, ctx.eps,
                                                               [True, True, True])
        return dy, None, dw, db, None


def kron_linear(state: KronLinearState, factors, weight, bias, drop_p=0.0, training=False, seed=None, weight_key=None,
                want_stats=False):
    """Public functional form (used by the modules below and by the tests).  `weight_key` identifies the weight's
    CONTENT for the packed-copy cache when `weight` is a temporary derived from a parameter (defaults to the
    tensor's own storage pointer + version counter)."""
    for f in factors:
        if not f.is_cuda:
            raise RuntimeError("Kronecker fusion runs on CUDA tensors only (no CPU fallback)")
        if f.dtype != torch.float32:
            raise RuntimeError(f"Kronecker fusion computes from fp32 factors; got {f.dtype}")
    if seed is None:
        # one 64-bit word per call from torch's CUDA generator: no host sync, reproducible under torch.manual_seed, and
        # graph-safe (a captured step re-draws it on every replay; the kernels read it from device memory)
        seed = (torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=factors[0].device)
                if (training and drop_p > 0) else 0)
    return _KronLinearFn.apply(state, weight, bias, drop_p, training, seed, weight_key, bool(want_stats), *factors)


# --------------------------------------------------------------------------- #
# modules
# --------------------------------------------------------------------------- #
class _GatedKronFusion(nn.Module):
    """Shared machinery of the three fusion modules.  Sub-modules are created in the reference's
    order (h, z, o per branch; post_fusion_dropout; encoder1; encoder2) so that state_dict keys and
    the global-RNG draws match fusion.py."""

    def _build(self, dims_og, scales, gate_inputs, use_bilinear, skip, mmhid, dropout_rate, post_p, batchnorm, poly=False):
        dims = [d // s for d, s in zip(dims_og, scales)]                 # fusion.py:17 / :76
        self._dims = dims
        self._gate_inputs = gate_inputs                                  # per branch: which (a, b) feed linear_z
        for t, (d_og, d) in enumerate(zip(dims_og, dims), start=1):
            a, b = gate_inputs[t - 1]
            setattr(self, f"linear_h{t}", nn.Sequential(nn.Linear(d_og, d), nn.ReLU()))
            z = nn.Bilinear(dims_og[a], dims_og[b], d) if use_bilinear else nn.Sequential(nn.Linear(dims_og[a] + dims_og[b], d))
            setattr(self, f"linear_z{t}", z)
            setattr(self, f"linear_o{t}", nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Dropout(p=dropout_rate)))
        self.post_fusion_dropout = nn.Dropout(p=post_p)
        kk = 1
        for d in dims:
            kk *= d + 1
        skip_dim = (sum(dims) + len(dims)) if skip else 0
        norm = (lambda: [nn.BatchNorm1d(mmhid)]) if batchnorm else (lambda: [])
        self.encoder1 = nn.Sequential(nn.Linear(kk, mmhid), *norm(), nn.ReLU(), nn.Dropout(p=dropout_rate))
        if poly:     # stage-2 PolynomialFusion: a second Kronecker encoder, then the skip encoder (its fusion.py:31-34)
            self.encoder2 = nn.Sequential(nn.Linear(kk, mmhid), *norm(), nn.ReLU(), nn.Dropout(p=dropout_rate))
            self.encoder3 = nn.Sequential(nn.Linear(mmhid + skip_dim, mmhid), *norm(), nn.ReLU(), nn.Dropout(p=dropout_rate))
            self._kron2 = KronLinearState([mmhid, mmhid])
        else:
            self.encoder2 = nn.Sequential(nn.Linear(mmhid + skip_dim, mmhid), *norm(), nn.ReLU(), nn.Dropout(p=dropout_rate))
        init_max_weights(self)
        self._kron = KronLinearState(dims)
        # the nn.Bilinear gates are the same contraction without the appended 1
        self._zkron = {t: KronLinearState([dims_og[a], dims_og[b]]) for t, (a, b) in enumerate(gate_inputs, start=1)} \
            if use_bilinear else {}

    def invalidate_kron_caches(self):
        """Drop the packed TF32 weight copies: call after writing weights through `.data` (no version bump)."""
        for st in [getattr(self, "_kron", None), getattr(self, "_kron2", None), *getattr(self, "_zkron", {}).values()]:
            if st is not None:
                st.invalidate()

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate_kron_caches()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_kron_caches()
        return out

    def set_kron_path(self, path):
        """"auto" (tensor cores) or "simt" (exact fp32 CUDA cores) for every Kronecker contraction of the module."""
        self._kron.path = path
        if hasattr(self, "_kron2"):
            self._kron2.path = path
        for st in self._zkron.values():
            st.path = path

    def _bilinear_gate(self, t, zmod, va, vb):
        """nn.Bilinear(va, vb) = kron(va, vb) @ W.view(d, -1)^T + b (fusion.py:21,25,43,50) through the Kronecker
        kernels: the [d, da, db] weight is zero-padded to the append-1 layout [d, (da+1)(db+1)] (the border and
        corner columns multiply the appended 1s and are zero), so forward, dW and the input gradients all reuse
        K1/K2/K3; autograd slices dW back through the pad."""
        w = zmod.weight
        wpad = torch.nn.functional.pad(w, (0, 1, 0, 1)).flatten(1)
        return kron_linear(self._zkron[t], [va, vb], wpad, zmod.bias,
                           weight_key=("bilinear", w.data_ptr(), w._version, tuple(w.shape)))

    def _branch(self, t, vecs, gated):
        own = vecs[t - 1]
        if gated:
            a, b = self._gate_inputs[t - 1]
            h = getattr(self, f"linear_h{t}")(own)
            zmod = getattr(self, f"linear_z{t}")
            if self.use_bilinear:
                z = self._bilinear_gate(t, zmod, vecs[a].contiguous(), vecs[b].contiguous())
            else:
                z = zmod(torch.cat((vecs[a], vecs[b]), dim=1))
            return getattr(self, f"linear_o{t}")(torch.sigmoid(z) * h)
        return getattr(self, f"linear_o{t}")(own)

    def _branches(self, vecs, gates):
        """The gated branches (fusion.py:41-54 / :99-121) are independent chains of small kernels: branch 1 stays on the
        current stream, the others run beside it on side streams (and, by autograd's stream affinity, so do their backward
        chains).  The Python call order -- hence every dropout's Philox offset and the RNG order of the reference -- is
        unchanged, the results are bit-identical.  MML_FUSION_STREAMS=0 keeps the branches in line."""
        dev = vecs[0].device
        if dev.type != "cuda" or len(gates) < 2 or os.environ.get("MML_FUSION_STREAMS", "1") != "1":
            return [self._branch(t, vecs, g) for t, g in enumerate(gates, start=1)]
        sides = _BRANCH_STREAMS.get(dev)
        if sides is None:
            sides = _BRANCH_STREAMS[dev] = [torch.cuda.Stream(dev, priority=-1) for _ in range(2)]
        cur = torch.cuda.current_stream(dev)
        for s in sides[:len(gates) - 1]:
            s.wait_stream(cur)                    # the inputs are ready; nothing of branch 1 is queued yet
        outs = [self._branch(1, vecs, gates[0])]
        for t in range(2, len(gates) + 1):
            with torch.cuda.stream(sides[t - 2]):
                outs.append(self._branch(t, vecs, gates[t - 1]))
        for s, o in zip(sides, outs[1:]):
            cur.wait_stream(s)
            o.record_stream(cur)
        return outs

    def _encode(self, enc, state, factors):
        """Linear(kron) [-> BatchNorm1d -> ReLU] -> Dropout of one Kronecker encoder (fusion.py:29-30 / :94).  In training mode
        the BatchNorm batch statistics come out of the Kronecker kernel's epilogue and one finisher kernel does
        statistics + running-stat update + normalise + ReLU; eval mode, BatchNorm variants the
        finisher does not cover (no affine, cumulative momentum) and the exact-fp32 path keep the modules' own ops."""
        lin = enc[0]
        bn = enc[1] if isinstance(enc[1], nn.BatchNorm1d) else None
        fused = (bn is not None and self.training and bn.affine and bn.momentum is not None and state.path == "auto"
                 and isinstance(enc[2], nn.ReLU))
        y = kron_linear(state, factors, lin.weight, lin.bias, self.post_fusion_dropout.p, self.training, want_stats=fused)
        if not fused:
            return enc[1:](y)
        out = _BNReLUFn.apply(y, state.last_stats, bn.weight, bn.bias, bn)
        state.last_stats = None
        return enc[3:](out)

    def _fuse(self, outs):
        out = self._encode(self.encoder1, self._kron, outs)
        if self.skip:
            ones = outs[0].new_ones(outs[0].shape[0], 1)
            out = torch.cat([out] + [torch.cat((o, ones), 1) for o in outs], 1)
        return self.encoder2(out)


class BilinearFusion(_GatedKronFusion):
    def __init__(self, skip=1, use_bilinear=1, gate1=1, gate2=1, dim1=32, dim2=32,
                 scale_dim1=1, scale_dim2=1, mmhid=64, dropout_rate=0.25):
        super(BilinearFusion, self).__init__()
        self.skip = skip
        self.use_bilinear = use_bilinear
        self.gate1 = gate1
        self.gate2 = gate2
        self.relu = nn.ReLU(inplace=False)
        self._build([dim1, dim2], [scale_dim1, scale_dim2], [(0, 1), (0, 1)], use_bilinear, skip, mmhid,
                    dropout_rate, post_p=dropout_rate, batchnorm=True)

    def forward(self, vec1, vec2):
        vecs = [self.relu(vec1), self.relu(vec2)]                         # fusion.py:38-39
        return self._fuse(self._branches(vecs, (self.gate1, self.gate2)))


class PolynomialFusion(_GatedKronFusion):
    """`MIA 2023/stage2_unimodal_student/fusion.py:6-73`: BilinearFusion whose encoder1 output, with a 1 appended, is
    Kronecker-multiplied with ITSELF and contracted by a second encoder (4th-order fusion, :63-68), then the skip
    encoder.  Both Kronecker tensors stay on chip (K1/K2/K3, the second with both factors = encoder1's output).  As in
    the reference, `encoder2` is sized by (dim1+1)(dim2+1), so the module only runs when that equals (mmhid+1)^2."""

    def __init__(self, skip=1, use_bilinear=1, gate1=1, gate2=1, dim1=32, dim2=32,
                 scale_dim1=1, scale_dim2=1, mmhid=64, dropout_rate=0.25):
        super(PolynomialFusion, self).__init__()
        self.skip = skip
        self.use_bilinear = use_bilinear
        self.gate1 = gate1
        self.gate2 = gate2
        self.relu = nn.ReLU(inplace=False)
        self._build([dim1, dim2], [scale_dim1, scale_dim2], [(0, 1), (0, 1)], use_bilinear, skip, mmhid,
                    dropout_rate, post_p=dropout_rate, batchnorm=True, poly=True)

    def forward(self, vec1, vec2):
        vecs = [self.relu(vec1), self.relu(vec2)]
        o1, o2 = self._branches(vecs, (self.gate1, self.gate2))
        lin2 = self.encoder2[0]
        out12 = self._encode(self.encoder1, self._kron, [o1, o2])
        if lin2.weight.shape[1] != (out12.shape[1] + 1) ** 2:          # the reference fails in F.linear the same way
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({out12.shape[0]}x{(out12.shape[1] + 1) ** 2} "
                               f"and {lin2.weight.shape[1]}x{lin2.weight.shape[0]})")
        out12 = out12.contiguous()
        out = self._encode(self.encoder2, self._kron2, [out12, out12])
        if self.skip:
            ones = o1.new_ones(o1.shape[0], 1)
            out = torch.cat((out, torch.cat((o1, ones), 1), torch.cat((o2, ones), 1)), 1)
        return self.encoder3(out)


class _TrilinearFusion(_GatedKronFusion):
    _GRAPH_GATE = (1, 2)

    def __init__(self, skip=1, use_bilinear=1, gate1=1, gate2=1, gate3=1, dim1=32, dim2=32, dim3=32,
                 scale_dim1=1, scale_dim2=1, scale_dim3=1, mmhid=96, dropout_rate=0.25):
        super().__init__()
        self.skip = skip
        self.use_bilinear = use_bilinear
        self.gate1 = gate1
        self.gate2 = gate2
        self.gate3 = gate3
        # path gated by omic, graph gated by omic (A) or path (B), omic gated by path; post-fusion p is a fixed 0.25;
        # no input ReLU and no BatchNorm (fusion.py:93-95, :99-120)
        self._build([dim1, dim2, dim3], [scale_dim1, scale_dim2, scale_dim3], [(0, 2), self._GRAPH_GATE, (0, 2)],
                    use_bilinear, skip, mmhid, dropout_rate, post_p=0.25, batchnorm=False)

    def forward(self, vec1, vec2, vec3):
        vecs = [vec1, vec2, vec3]
        return self._fuse(self._branches(vecs, (self.gate1, self.gate2, self.gate3)))


class TrilinearFusion_A(_TrilinearFusion):
    """fusion.py:66-132 -- graph branch gated with (vec2, vec3)."""
    _GRAPH_GATE = (1, 2)


class TrilinearFusion_B(_TrilinearFusion):
    """fusion.py:135-201 -- graph branch gated with (vec2, vec1)."""
    _GRAPH_GATE = (1, 0)
