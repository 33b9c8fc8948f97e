"""Row-sharded ContrastMemory / CRDLoss over `torch.distributed` (one process per GPU, NCCL).

The reference keeps both memory banks on one GPU (`CL_utils/CRD_criterion.py:20-23`; nothing in
the reference is distributed).  Here rank r owns rows [r*rows_per, (r+1)*rows_per) of BOTH banks
and the batch is data-parallel (each rank holds B_local anchors and their contrast_idx).  Rows
never cross NVLink: INDICES and per-anchor partial results do.  One step is

  1. Embed heads on the LOCAL anchors, then all_gather of the embeddings v1|v2 and of idx|positives
     (KBs..MBs).  The heads' parameter gradients are summed across ranks by ONE all_reduce of the
     flattened parameter vector inside backward (`_SumGradAcrossRanks`).
  2. route           contrast_idx -> per-owner int32 local row ids (stable counting sort by owner).
       transport "peer"    : single-pass fixed-stride routing (`mml_shard_route_strided`) into a
                             SYMMETRIC-MEMORY buffer; owners PULL their slots over NVLink from inside
                             the gather kernel (`mml_crd_fused_loss_grad_peer`): the all_to_all is
                             fused into K4 as peer loads (4 B of index per 1 KB of local row traffic),
                             no host sync, no staging copy.  Only the per-slot counts (36 KB/peer)
                             go through a fixed-size all_to_all, which doubles as the "routing is
                             complete everywhere" barrier.
       transport "alltoall": compacting two-pass routing (`mml_shard_count/scatter`) + a variable
                             size NCCL all_to_all (one host sync for the split sizes).  Used when
                             peer mapping is unavailable, and on CPU/gloo in the tests.
  3. local kernel    every rank scores ALL global anchors against the rows IT owns:
                     partial log-term sums, partial dL/dv.
  4. reduce          "peer": the anchors' home ranks pull-reduce their rows of every rank's partial buffer
                     (`mml_symm_pull_reduce`); "alltoall": one NCCL all_reduce of [sums | dL/dv1 | dL/dv2].
  5. owner update    each rank updates the rows of idx it owns (`mml_crd_memory_update` with
                     row_begin/row_end); gather-before-update ordering is stream order on each rank.

Per-rank gather traffic is (K+1)/world columns for every global anchor = the single-GPU unit of
work when the global batch grows with the world size (weak scaling), and there is no bulk row
exchange at all.  Result == the single-GPU module on the same global batch (integer routing exact;
float sums re-associated across ranks, tolerance 1e-4).

The compute steps go through a small backend object so the collective choreography can be tested
on CPU (gloo, world_size 2) with the oracle standing in for the CUDA kernels -- tests only; the
default backend is CUDA-only and fails loudly without the library.
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist
from torch import nn
from torch.autograd.function import once_differentiable

from . import _cabi
from . import crd as _crd
from .crd import AliasMethod, ContrastLoss, Embed


class CudaBackend:
    """The product backend: every compute step is a libmml_b200.so kernel."""

    def route(self, cidx, rows_per, world):
        """-> (counts int64 [world, B], ids int32 [B*cols] grouped by owner, (anchor, column)-stable)."""
        lib = _cabi.lib()
        B, cols = cidx.shape
        dev = cidx.device
        st = _cabi.cur_stream(dev)
        chunk = 512                                         # columns per warp: enough warps to fill the GPU
        chunks = (cols + chunk - 1) // chunk
        fine = torch.empty(world, B, chunks, dtype=torch.int64, device=dev)
        _cabi.check(lib.mml_shard_count(_cabi.dptr(cidx, torch.int64), B, cols, chunk, rows_per, world, _cabi.dptr(fine), st),
                    "mml_shard_count")
        flat = fine.reshape(-1)
        offsets = (flat.cumsum(0) - flat).contiguous()
        ids = torch.empty(B * cols, dtype=torch.int32, device=dev)
        _cabi.check(lib.mml_shard_scatter(_cabi.dptr(cidx, torch.int64), B, cols, chunk, rows_per, world,
                                          _cabi.dptr(offsets), _cabi.dptr(ids), st), "mml_shard_scatter")
        return fine.sum(2), ids

    def stats(self, bank1, bank2, v1, v2, ids, seg_ptr, T, cols):
        """-> sums[4] with raw exp sums in [2], [3] (first-step Z, CRD_criterion.py:52-59)."""
        return _crd.crd_scores(bank1, bank2, v1, v2, ids, T, seg_ptr=seg_ptr, cols=cols, want_out=False, want_sums=True)[2]

    def fused(self, bank1, bank2, v1, v2, ids, seg_ptr, pos_flag, T, Z, n_data, nce_k, batch):
        """-> (sums[4], g1[B,D], g2[B,D]) partial over the rows of this shard."""
        _, sums, g1, g2, _, _ = _crd.crd_fused_loss_grad(bank1, bank2, v1, v2, ids, T, Z, n_data, nce_k, seg_ptr=seg_ptr,
                                                        pos_flag=pos_flag, batch_norm=batch, want_sums=True, cols=nce_k + 1)
        return sums, g1, g2

    def update(self, bank1, bank2, v1, v2, y, momentum, row_begin, row_end):
        _crd.crd_memory_update(bank1, bank2, v1, v2, y, momentum, row_begin, row_end)

    # ---- peer (NVLink pull) transport ----
    def route_strided(self, cidx, rows_per, world, chunk, counts, ids):
        lib = _cabi.lib()
        B, cols = cidx.shape
        _cabi.check(lib.mml_shard_route_strided(_cabi.dptr(cidx, torch.int64), B, cols, chunk, rows_per, world,
                                                _cabi.dptr(counts), _cabi.dptr(ids), _cabi.cur_stream(cidx.device)),
                    "mml_shard_route_strided")

    def stats_peer(self, bank1, bank2, v1, v2, px, T):
        lib = _cabi.lib()
        Bg, D = v1.shape
        dev = v1.device
        nws = lib.mml_crd_peer_workspace_bytes(Bg, px.chunks, D)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        sums = torch.empty(4, dtype=torch.float32, device=dev)
        _cabi.check(lib.mml_crd_scores_peer(_cabi.dptr(bank1), _cabi.dptr(bank2), bank1.shape[0], D, _cabi.dptr(v1),
                                            _cabi.dptr(v2), px.ids_ptrs, px.cnt_ptrs, px.world, px.B_local,
                                            px.chunks, px.chunk, float(T), _cabi.dptr(sums), _cabi.dptr(ws), nws,
                                            _cabi.cur_stream(dev)), "mml_crd_scores_peer")
        return sums

    def fused_peer(self, bank1, bank2, v1, v2, px, pos_flag, T, Z, n_data, nce_k, batch, g1, g2, sums):
        """Partial sums / gradients over this shard's rows, written into the caller's (symmetric) buffers."""
        lib = _cabi.lib()
        Bg, D = v1.shape
        dev = v1.device
        nws = lib.mml_crd_peer_workspace_bytes(Bg, px.chunks, D)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        if _crd.KERNEL_TIMER is not None:
            _crd.KERNEL_TIMER.start("crd_fused_loss_grad", dev)
        rc = lib.mml_crd_fused_loss_grad_peer(
            _cabi.dptr(bank1), _cabi.dptr(bank2), bank1.shape[0], D, _cabi.dptr(v1), _cabi.dptr(v2), px.ids_ptrs,
            px.cnt_ptrs, px.world, px.B_local, px.chunks, px.chunk, _cabi.dptr(pos_flag), float(T),
            _cabi.dptr(Z), int(n_data), int(nce_k), int(batch), _cabi.dptr(sums), _cabi.dptr(g1), _cabi.dptr(g2),
            _cabi.dptr(ws), nws, _cabi.cur_stream(dev))
        if _crd.KERNEL_TIMER is not None:
            _crd.KERNEL_TIMER.stop("crd_fused_loss_grad", dev)
        _cabi.check(rc, "mml_crd_fused_loss_grad_peer")


class PeerExchange:
    """One symmetric-memory arena per (B_local, cols, D) geometry.  Every rank allocates the same layout; peers'
    arenas are reachable through CUDA peer mappings (`buffer_ptrs`).  Regions:
      ids  int32 [2][world][B_l][chunks][chunk]   routed local row ids, block o is read by owner o; two buffers, so that the
                                               next step's indices can be routed while this step's are still being read
      cnt  int32 [2][world][B_l][chunks]          entries used in each slot
      V    float [2][2][B_g][D]                all-gathered embeddings v1|v2, double-buffered by step parity
      YP   int64 [2][2][B_g]                   all-gathered anchor ids | positive rows, double-buffered
      P    float [2][B_g][D] + tail[16]        this rank's partial dL/dv1|dL/dv2 and its 4 partial sums"""

    def __init__(self, mem, B_local, cols, D, device, grad_floats=0):
        import ctypes

        import torch.distributed._symmetric_memory as symm
        self.world, self.rank, self.B_local, self.D = mem.world, mem.rank, B_local, D
        self.Bg = Bg = self.world * B_local
        self.chunk = min(2048, (cols + 31) // 32 * 32)
        self.chunks = (cols + self.chunk - 1) // self.chunk
        al = lambda n: (n + 255) // 256 * 256
        block = B_local * self.chunks * self.chunk * 4
        cblock = B_local * self.chunks * 4
        self.off_ids = [0, al(self.world * block)]
        self.off_cnt = [al(self.off_ids[1] + self.world * block), 0]
        self.off_cnt[1] = al(self.off_cnt[0] + self.world * cblock)
        self.off_V = al(self.off_cnt[1] + self.world * cblock)
        self.off_YP = al(self.off_V + 2 * 2 * Bg * D * 4)
        self.off_P = al(self.off_YP + 2 * 2 * Bg * 8)
        self.off_tail = self.off_P + 2 * Bg * D * 4
        self.grad_pad = (int(grad_floats) + 3) // 4 * 4                 # G float [world][grad_pad]: every rank's head gradients
        self.off_G = al(self.off_tail + 64)
        total = al(self.off_G + self.world * self.grad_pad * 4 + 64)
        group = mem.group if mem.group is not None else dist.group.WORLD
        if hasattr(symm, "enable_symm_mem_for_group"):
            try:
                symm.enable_symm_mem_for_group(group.group_name)
            except Exception:
                pass
        self.raw = symm.empty(total, dtype=torch.uint8, device=device)
        self.raw.zero_()
        self.handle = symm.rendezvous(self.raw, group)
        bases = [int(p) for p in self.handle.buffer_ptrs]
        P = ctypes.c_void_p
        self.base_ptrs = (P * self.world)(*bases)
        view = lambda off, nbytes, dt, shape: self.raw[off:off + nbytes].view(dt).view(shape)
        self._route_bufs = [
            ((P * self.world)(*[b + self.off_ids[k] + self.rank * block for b in bases]),
             (P * self.world)(*[b + self.off_cnt[k] + self.rank * cblock for b in bases]),
             view(self.off_ids[k], self.world * block, torch.int32, (-1,)),
             view(self.off_cnt[k], self.world * cblock, torch.int32, (self.world, B_local, self.chunks)))
            for k in (0, 1)]
        self.use_route_buffer(0)
        self.V = view(self.off_V, 2 * 2 * Bg * D * 4, torch.float32, (2, 2, Bg, D))
        self.YP = view(self.off_YP, 2 * 2 * Bg * 8, torch.int64, (2, 2, Bg))
        self.P = view(self.off_P, 2 * Bg * D * 4, torch.float32, (2, Bg, D))
        self.tail = view(self.off_tail, 16, torch.float32, (4,))
        self.G = (view(self.off_G, self.world * self.grad_pad * 4, torch.float32, (self.world, self.grad_pad))
                  if self.grad_pad else None)
        self.step = 0
        self._c = ctypes

    def use_route_buffer(self, k):
        """Make routing buffer k the one the routing / gather calls of the current step work on."""
        self.route_buf = k
        self.ids_ptrs, self.cnt_ptrs, self.ids, self.counts = self._route_bufs[k]

    def push(self, v1, v2, idx, pos, buf):
        """all_gather by NVLink stores: my slices land in every rank's V / YP buffers of parity `buf`."""
        c, Bl, Bg, D, r = self._c, self.B_local, self.Bg, self.D, self.rank
        vb = self.off_V + buf * (2 * Bg * D * 4)
        yb = self.off_YP + buf * (2 * Bg * 8)
        srcs = (c.c_void_p * 4)(v1.data_ptr(), v2.data_ptr(), idx.data_ptr(), pos.data_ptr())
        offs = (c.c_int64 * 4)(vb + r * Bl * D * 4, vb + (Bg + r * Bl) * D * 4, yb + r * Bl * 8, yb + (Bg + r * Bl) * 8)
        nbytes = (c.c_int64 * 4)(Bl * D * 4, Bl * D * 4, Bl * 8, Bl * 8)
        _cabi.check(_cabi.lib().mml_symm_push(self.base_ptrs, self.world, srcs, offs, nbytes, 4,
                                              _cabi.cur_stream(v1.device)), "mml_symm_push")

    def allreduce_sum(self, g):
        """Sum of a small flat fp32 vector over the ranks without NCCL: every rank STORES its vector into slice `rank` of
        every peer's G region (NVLink), one barrier, then each rank adds the `world` slices of its OWN region in rank
        order -- the same operands in the same order everywhere, so replicated parameters stay bit-identical.  Reuse
        across steps is ordered by the two barriers of the next step's forward."""
        n = g.numel()
        if self.G is None or n > self.grad_pad:
            raise RuntimeError("PeerExchange was built without room for this gradient vector")
        c = self._c
        src = g if n == self.grad_pad else torch.cat((g.reshape(-1), g.new_zeros(self.grad_pad - n)))
        srcs = (c.c_void_p * 1)(src.data_ptr())
        offs = (c.c_int64 * 1)(self.off_G + self.rank * self.grad_pad * 4)
        nbytes = (c.c_int64 * 1)(self.grad_pad * 4)
        _cabi.check(_cabi.lib().mml_symm_push(self.base_ptrs, self.world, srcs, offs, nbytes, 1,
                                              _cabi.cur_stream(g.device)), "mml_symm_push")
        self.handle.barrier(channel=2)
        return self.G[:, :n].sum(0).view_as(g)

    def pull_reduce(self, g1, g2, tail):
        """reduce_scatter by NVLink loads: my anchors' rows summed over every rank's partial buffer."""
        _cabi.check(_cabi.lib().mml_symm_pull_reduce(
            self.base_ptrs, self.world, self.off_P, self.Bg, self.rank * self.B_local, self.B_local, self.D,
            _cabi.dptr(g1), _cabi.dptr(g2), self.off_tail, 4, _cabi.dptr(tail), _cabi.cur_stream(g1.device)),
            "mml_symm_pull_reduce")


class _SumGradAcrossRanks(torch.autograd.Function):
    """Identity on a list of parameters in forward; backward concatenates their gradients (ONE kernel), all-reduces (SUM)
    the flat vector with one collective and hands the slices back.  (A flat `torch.cat` of the parameters sliced into
    views costs a copy in forward and, in backward, a fill + copy + add per parameter: 37 us of 1-2 us kernels per step in
    the timeline of the 2-GPU step, `profiles/r2_sharded_timeline_g2.txt`.)"""

    @staticmethod
    def forward(ctx, group, reducer, *params):
        ctx.group, ctx.reducer = group, reducer
        ctx.shapes = [p.shape for p in params]
        return tuple(p.view_as(p) for p in params)

    @staticmethod
    def backward(ctx, *grads):
        ref = next(g for g in grads if g is not None)
        parts = [(g if g is not None else ref.new_zeros(shp)).reshape(-1) for g, shp in zip(grads, ctx.shapes)]
        flat = torch.cat(parts)
        if ctx.reducer is not None:          # peer transport: symmetric-memory exchange, no NCCL on the step
            flat = ctx.reducer(flat)
        else:
            dist.all_reduce(flat, group=ctx.group)
        outs, off = [], 0
        for shp in ctx.shapes:
            n = math.prod(shp)
            outs.append(flat[off:off + n].view(shp))
            off += n
        return (None, None, *outs)


class _AllGatherRows(torch.autograd.Function):
    """all_gather along dim 0.  Backward takes the local slice: every rank computes the SAME full
    gradient (the downstream computation is replicated), so no reduction is needed."""

    @staticmethod
    def forward(ctx, x, group):
        world = dist.get_world_size(group)
        ctx.rank, ctx.n = dist.get_rank(group), x.shape[0]
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return g[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n], None


def _all_gather(x, group):
    world = dist.get_world_size(group)
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


class _ShardedFusedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, V1, V2, mem, Y, routed, n_data):
        loss, g1, g2 = mem._sharded_step(V1.detach().contiguous(), V2.detach().contiguous(), Y, routed, n_data)
        ctx.save_for_backward(g1, g2)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loss):
        g1, g2 = ctx.saved_tensors
        return grad_loss * g1, grad_loss * g2, None, None, None, None


class _ShardedPeerFn(torch.autograd.Function):
    """Sharded step with every exchange done over peer-mapped memory (see module docstring, transport "peer")."""

    @staticmethod
    def forward(ctx, v1, v2, mem, idx, cidx, n_data):
        loss, g1, g2 = mem._peer_step(v1.detach().contiguous(), v2.detach().contiguous(), idx, cidx, n_data)
        ctx.save_for_backward(g1, g2)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loss):
        g1, g2 = ctx.saved_tensors
        return grad_loss * g1, grad_loss * g2, None, None, None, None


class ShardedContrastMemory(nn.Module):
    """ContrastMemory(inputSize, outputSize, K, T, momentum) with its rows split over the ranks of `group`.
    Buffers: `params` (replicated, same layout as the reference), `memory_v1` / `memory_v2` = the LOCAL
    row block [rows_local, inputSize] (`row_begin`, `row_end` say which)."""

    def __init__(self, inputSize, outputSize, K, T=0.07, momentum=0.5, group=None, device=None, backend=None,
                 init="rand"):
        super().__init__()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nLem = outputSize
        self.K = K
        self.rows_per = (outputSize + self.world - 1) // self.world
        self.row_begin = min(outputSize, self.rank * self.rows_per)
        self.row_end = min(outputSize, self.row_begin + self.rows_per)
        self.backend = backend if backend is not None else CudaBackend()
        rows_local = self.row_end - self.row_begin
        self.register_buffer("params", torch.tensor([K, T, -1, -1, momentum]))
        stdv = 1. / math.sqrt(inputSize / 3)
        dev = device if device is not None else "cpu"
        if init == "rand":      # same distribution as CRD_criterion.py:21-23, drawn per shard on its device
            m1 = torch.rand(rows_local, inputSize, device=dev).mul_(2 * stdv).add_(-stdv)
            m2 = torch.rand(rows_local, inputSize, device=dev).mul_(2 * stdv).add_(-stdv)
        else:
            m1 = torch.empty(rows_local, inputSize, device=dev)
            m2 = torch.empty(rows_local, inputSize, device=dev)
        self.register_buffer("memory_v1", m1)
        self.register_buffer("memory_v2", m2)
        if device is not None:
            self.params = self.params.to(device)
        self.multinomial = None          # built lazily: only the idx=None branch samples
        self._peer = {}                  # (B_local, cols, device) -> PeerExchange
        self._side_stream = None         # routing runs here, under the Embed heads
        self._route_side = None          # set by route_ahead, consumed by _peer_step
        self._prefetched = None          # (data_ptr, arena, routing buffer, stream) of a prefetched NEXT batch
        self._routed_ptr = None          # data_ptr of the contrast_idx already routed for the current step
        self._next_cidx = None           # next batch handed to forward(), routed right after this step's gather
        self.grad_floats = 0             # room for the data-parallel heads' flattened gradient in each arena
        p = torch.tensor([K, T, -1, -1, momentum])
        self._K, self._T, self._momentum = int(p[0].item()), p[1].item(), p[4].item()
        self._z_ready = False

    # -- host-side scalar cache, as ContrastMemory: K, T, momentum and the Z state follow `params` wherever it is loaded from
    def _refresh_scalars(self):
        p = self.params.detach().cpu()
        self._K = int(p[0].item())
        self._T = p[1].item()
        self._momentum = p[4].item()
        self._z_ready = bool(p[2].item() > 0 and p[3].item() > 0)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._refresh_scalars()

    # ---- (de)sharding of the reference's state-dict layout ----
    def load_full_banks(self, memory_v1, memory_v2, params=None):
        """Take this rank's row block out of full [n, D] banks (e.g. a single-GPU checkpoint)."""
        self.memory_v1.copy_(memory_v1[self.row_begin:self.row_end])
        self.memory_v2.copy_(memory_v2[self.row_begin:self.row_end])
        if params is not None:
            self.params.copy_(params)
            self._refresh_scalars()

    def gather_full_banks(self):
        """-> (memory_v1, memory_v2) as full [n, D] tensors on every rank (the reference's layout)."""
        outs = []
        for bank in (self.memory_v1, self.memory_v2):
            pad = torch.zeros(self.rows_per, bank.shape[1], dtype=bank.dtype, device=bank.device)
            pad[:bank.shape[0]] = bank
            outs.append(_all_gather(pad, self.group)[:self.nLem])
        return tuple(outs)

    # ---- one sharded step on replicated V1/V2 (global batch) ----
    def exchange_begin(self, cidx_local):
        """Route this rank's contrast_idx to the owners and start the size exchange.  The one host sync of
        the step (all_to_all needs split sizes on the host) is deferred to `exchange_finish`, so the caller
        can queue independent GPU work (all_gathers, Embed heads) in between."""
        cols = self._K + 1
        if cidx_local.dim() != 2 or cidx_local.shape[1] != cols:
            raise RuntimeError(f"contrast_idx must be [B, nce_k+1 = {cols}], got {tuple(cidx_local.shape)}")   # :42
        counts, ids = self.backend.route(cidx_local, self.rows_per, self.world)      # [world, B], [B*cols]
        recv_counts = torch.empty_like(counts)                                       # [src, B]
        dist.all_to_all_single(recv_counts, counts, group=self.group)
        sizes = torch.stack((counts.sum(1), recv_counts.sum(1)))
        if sizes.is_cuda:
            host = torch.empty(sizes.shape, dtype=sizes.dtype, pin_memory=True)
            host.copy_(sizes, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:
            host, ev = sizes, None
        return ids, recv_counts, host, ev

    def peer_arena(self, B_local, cols, D, device, grad_floats=0):
        key = (B_local, cols, D, device)
        px = self._peer.get(key)
        if px is None:
            px = self._peer[key] = PeerExchange(self, B_local, cols, D, device, grad_floats=max(grad_floats, self.grad_floats))
        return px

    def route_ahead(self, cidx, D):
        """Start the routing of this step's contrast_idx on a side stream.  Routing needs nothing but the indices, so it
        runs UNDER the Embed heads (small GEMMs that leave most SMs idle) instead of in front of the gather; `_peer_step`
        joins the side stream before it publishes the slots.  The fork happens after everything already queued on the
        current stream -- in particular after the previous step's post-gather barrier, so no peer is still reading the
        slots that get overwritten.  Works inside a CUDA-graph capture (fork / join become graph dependencies)."""
        cols = self._K + 1
        if cidx.dim() != 2 or cidx.shape[1] != cols or not cidx.is_cuda:
            return
        dev = cidx.device
        px = self.peer_arena(cidx.shape[0], cols, D, dev)
        if self._claim_prefetched(px, cidx):
            return                                  # routed during the previous step (`prefetch_routing`)
        self._prefetched = None                     # whatever was prefetched is not this batch: its buffer is reused below
        px.use_route_buffer(1 - px.route_buf)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(dev)
        side = self._side_stream
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self.backend.route_strided(cidx, self.rows_per, self.world, px.chunk, px.counts, px.ids)
        cidx.record_stream(side)
        self._route_side = side
        self._routed_ptr = cidx.data_ptr()

    def _claim_prefetched(self, px, cidx):
        """True when `cidx` is the tensor whose routing `prefetch_routing` started during the previous step: switch to
        the buffer it was routed into and make the current stream wait for that routing."""
        pre = self._prefetched
        if pre is None or pre[0] != cidx.data_ptr() or pre[1] is not px:
            return False
        self._prefetched = None
        px.use_route_buffer(pre[2])
        if not torch.cuda.is_current_stream_capturing():      # replayed graphs are ordered by their launches: the routing was
            torch.cuda.current_stream(cidx.device).wait_stream(pre[3])      # joined at the end of the previous captured step
        self._routed_ptr = cidx.data_ptr()
        return True

    def prefetch_routing(self, next_cidx, D):
        """Route the NEXT step's contrast_idx now, on a side stream, into the routing buffer this step does not use.
        Called right after this step's gather has been enqueued: the routing kernel then runs under the post-gather
        barrier (waiting for the slowest rank), the pull-reduce and the Embed heads' backward -- ~300 us of kernels that
        leave most SMs idle -- instead of in front of the next gather (105 us, `profiles/r2_sharded_timeline_g2.txt`).
        No peer reads the other buffer before the next step's pre-gather barrier, which this rank reaches only after
        joining the side stream.  Inside a CUDA-graph capture the join is deferred to the end of the captured step
        (`graphed.defer_to_end_of_step`)."""
        cols = self._K + 1
        if next_cidx is None or next_cidx.dim() != 2 or next_cidx.shape[1] != cols or not next_cidx.is_cuda:
            return
        dev = next_cidx.device
        px = self.peer_arena(next_cidx.shape[0], cols, D, dev)
        other = 1 - px.route_buf
        _, _, ids, counts = px._route_bufs[other]
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(dev)
        side = self._side_stream
        cur = torch.cuda.current_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.backend.route_strided(next_cidx, self.rows_per, self.world, px.chunk, counts, ids)
        next_cidx.record_stream(side)
        self._prefetched = (next_cidx.data_ptr(), px, other, side)
        from . import graphed as _graphed
        _graphed.defer_to_end_of_step(lambda: torch.cuda.current_stream(dev).wait_stream(side))

    def _peer_step(self, v1, v2, idx, cidx, n_data):
        """Local embeddings in, loss (global batch) + local dL/dv out.  Exchanges: routed ids / counts are PULLED by
        the owners inside K4; embeddings and anchor ids are PUSHED into every rank's buffers; partial gradients are
        PULL-reduced by the anchors' home ranks.  Two symmetric-memory barriers per step, no NCCL, no host sync."""
        be = self.backend
        cols = self._K + 1
        if cidx.dim() != 2 or cidx.shape[1] != cols:
            raise RuntimeError(f"contrast_idx must be [B, nce_k+1 = {cols}], got {tuple(cidx.shape)}")   # :42
        Bl, D = v1.shape
        px = self.peer_arena(Bl, cols, D, v1.device)
        buf = px.step & 1
        px.step += 1
        if self._route_side is not None:          # routed ahead of the Embed heads on a side stream (route_ahead): join it
            torch.cuda.current_stream(v1.device).wait_stream(self._route_side)
            self._route_side = None
        elif self._routed_ptr != cidx.data_ptr() and not self._claim_prefetched(px, cidx):
            self._prefetched = None
            px.use_route_buffer(1 - px.route_buf)
            be.route_strided(cidx, self.rows_per, self.world, px.chunk, px.counts, px.ids)
        self._routed_ptr = None
        px.push(v1, v2, idx.contiguous(), cidx[:, 0].contiguous(), buf)
        px.handle.barrier(channel=0)                  # everyone's slots, counts and slices are in place
        V1, V2, Y, pos_rows = px.V[buf, 0], px.V[buf, 1], px.YP[buf, 0], px.YP[buf, 1]
        pos_flag = ((pos_rows >= self.row_begin) & (pos_rows < self.row_end)).to(torch.uint8)
        Bg = px.Bg
        if not self._z_ready:                                                       # CRD_criterion.py:52-59
            sums = be.stats_peer(self.memory_v1, self.memory_v2, V1, V2, px, self._T)
            dist.all_reduce(sums, group=self.group)
            scale = float(self.nLem) / (float(Bg) * cols)
            z = self.params[2:4]
            self.params[2:4] = torch.where(z < 0, sums[2:4] * scale, z)
            if self.rank == 0:
                print("normalization constant Z_v1 is set to {:.1f}".format(self.params[2].item()))
                print("normalization constant Z_v2 is set to {:.1f}".format(self.params[3].item()))
            self._z_ready = True
        be.fused_peer(self.memory_v1, self.memory_v2, V1, V2, px, pos_flag, self._T, self.params[2:4], n_data,
                      self._K, Bg, px.P[0], px.P[1], px.tail)
        if self._next_cidx is not None:               # the caller knows the next batch already: route it under this step's tail
            nxt, self._next_cidx = self._next_cidx, None
            self.prefetch_routing(nxt, D)
        px.handle.barrier(channel=1)                  # every rank's partials are complete
        g1 = torch.empty(Bl, D, dtype=torch.float32, device=v1.device)
        g2 = torch.empty(Bl, D, dtype=torch.float32, device=v1.device)
        tail = torch.empty(4, dtype=torch.float32, device=v1.device)
        px.pull_reduce(g1, g2, tail)
        loss = (-(tail[0] + tail[1]) / Bg).reshape(1)
        with torch.no_grad():                                                       # :66-79, owner applies
            be.update(self.memory_v1, self.memory_v2, V1, V2, Y, self._momentum, self.row_begin, self.row_end)
        return loss, g1, g2

    def exchange_finish(self, handle):
        """-> (ids int32 [nnz], seg_ptr int64 [B_global+1]): every rank's requests for MY rows, global anchor order."""
        ids, recv_counts, host, ev = handle
        if ev is not None:
            ev.synchronize()
        send_l, recv_l = host[0].tolist(), host[1].tolist()
        recv_ids = torch.empty(int(sum(recv_l)), dtype=torch.int32, device=ids.device)
        dist.all_to_all_single(recv_ids, ids, output_split_sizes=recv_l, input_split_sizes=send_l, group=self.group)
        flat = recv_counts.reshape(-1)
        seg_ptr = torch.zeros(flat.numel() + 1, dtype=torch.int64, device=flat.device)
        seg_ptr[1:] = flat.cumsum(0)
        return recv_ids, seg_ptr

    def _sharded_step(self, V1, V2, Y, routed, n_data):
        be, group = self.backend, self.group
        Bg, D = V1.shape
        cols = self._K + 1
        ids, seg_ptr, pos_rows = routed
        pos_flag = ((pos_rows >= self.row_begin) & (pos_rows < self.row_end)).to(torch.uint8)
        if not self._z_ready:                                                       # CRD_criterion.py:52-59
            sums = be.stats(self.memory_v1, self.memory_v2, V1, V2, ids, seg_ptr, self._T, cols).clone()
            dist.all_reduce(sums, group=group)
            scale = float(self.nLem) / (float(Bg) * cols)
            z = self.params[2:4]
            new_z = sums[2:4] * scale
            self.params[2:4] = torch.where(z < 0, new_z, z)
            if self.rank == 0:
                print("normalization constant Z_v1 is set to {:.1f}".format(self.params[2].item()))
                print("normalization constant Z_v2 is set to {:.1f}".format(self.params[3].item()))
            self._z_ready = True
        sums, g1, g2 = be.fused(self.memory_v1, self.memory_v2, V1, V2, ids, seg_ptr, pos_flag, self._T,
                                self.params[2:4], n_data, self._K, Bg)
        packed = torch.cat((sums.reshape(-1)[:2], g1.reshape(-1), g2.reshape(-1)))
        dist.all_reduce(packed, group=group)
        loss = (-(packed[0] + packed[1]) / Bg).reshape(1)
        g1 = packed[2:2 + Bg * D].view(Bg, D)
        g2 = packed[2 + Bg * D:].view(Bg, D)
        with torch.no_grad():                                                       # :66-79, owner applies
            be.update(self.memory_v1, self.memory_v2, V1, V2, Y, self._momentum, self.row_begin, self.row_end)
        return loss, g1, g2

    def fused_nce_loss(self, V1, V2, Y, routed, n_data):
        return _ShardedFusedFn.apply(V1, V2, self, Y, routed, n_data)


class ShardedCRDLoss(nn.Module):
    """CRDLoss(opt) over a row-sharded bank.  `forward(f_s, f_t, idx, contrast_idx)` takes this rank's LOCAL
    batch and returns the loss of the GLOBAL batch (identical on every rank); `loss.backward()` leaves the
    local rows' gradient in f_s.grad and the GLOBAL-batch gradient in the (replicated) Embed parameters.
    transport: "peer" (NVLink pull through symmetric memory; default on CUDA), "alltoall", or "auto"."""

    def __init__(self, opt, group=None, device=None, backend=None, transport="auto"):
        super().__init__()
        self.group = group
        self.embed_s = Embed(opt.s_dim, opt.feat_dim)
        self.embed_t = Embed(opt.t_dim, opt.feat_dim)
        if device is not None:
            self.embed_s.to(device)
            self.embed_t.to(device)
        if dist.get_world_size(group) > 1:          # replicated heads must start identical
            for p in list(self.embed_s.parameters()) + list(self.embed_t.parameters()):
                dist.broadcast(p.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.contrast = ShardedContrastMemory(opt.feat_dim, opt.n_data, opt.nce_k, opt.nce_t, opt.nce_m, group=group,
                                              device=device, backend=backend)
        self.criterion_t = ContrastLoss(opt.n_data)
        self.criterion_s = ContrastLoss(opt.n_data)
        self.transport = transport
        # "symm": the heads' gradient is summed over peer-mapped memory (no NCCL anywhere on the step); "nccl": all_reduce
        self.heads_reduce = os.environ.get("MML_HEADS_REDUCE", "symm")
        self._feat_dim = opt.feat_dim

    def _pick_transport(self, device):
        if self.transport == "auto":
            self.transport = "peer" if (device.type == "cuda" and isinstance(self.contrast.backend, CudaBackend)) else "alltoall"
        return self.transport

    def _heads(self, f_s, f_t, reducer=None):
        """Embed heads on the local anchors with parameters routed through `_SumGradAcrossRanks`."""
        named = [(k, p) for k, p in self.named_parameters()]
        routed = _SumGradAcrossRanks.apply(self.group, reducer, *[p for _, p in named])
        views = {k: v for (k, _), v in zip(named, routed)}
        ps = {k[len("embed_s."):]: v for k, v in views.items() if k.startswith("embed_s.")}
        pt = {k[len("embed_t."):]: v for k, v in views.items() if k.startswith("embed_t.")}
        return _crd.embed_pair(lambda x: torch.func.functional_call(self.embed_s, ps, (x,)), f_s,
                               lambda x: torch.func.functional_call(self.embed_t, pt, (x,)), f_t)

    def forward(self, f_s, f_t, idx, contrast_idx=None, next_contrast_idx=None):
        """next_contrast_idx (extension, peer transport): the contrast_idx of the NEXT call, if the caller already holds it on
        the device (a prefetching loader does) -- its routing then runs under this step's tail instead of in front of the
        next gather.  The next call must pass that same tensor as `contrast_idx`."""
        g, mem = self.group, self.contrast
        from . import graphed as _graphed
        _graphed.run_deferred()                     # joins left over by a previous step outside a GraphedTrainStep
        if contrast_idx is None:                    # CRD_criterion.py:37-39 on the local anchors
            if mem.multinomial is None:
                mem.multinomial = AliasMethod(torch.ones(mem.nLem))
                mem.multinomial.cuda(f_s.device)
            B = idx.shape[0]
            contrast_idx = mem.multinomial.draw(B * (mem.K + 1), y=idx, cols=mem.K + 1).view(B, -1)
        contrast_idx = contrast_idx.contiguous()
        transport = self._pick_transport(f_s.device)
        if transport == "peer" and idx.shape[0] % 2:
            transport = "alltoall"                  # 16-byte push granularity needs an even local batch
        if transport == "peer":
            try:
                reducer = None
                if self.heads_reduce == "symm":
                    mem.grad_floats = sum(p.numel() for p in self.parameters())
                    px = mem.peer_arena(idx.shape[0], mem._K + 1, self._feat_dim, f_s.device)
                    reducer = px.allreduce_sum
                mem._next_cidx = next_contrast_idx.contiguous() if next_contrast_idx is not None else None
                if os.environ.get("MML_ROUTE_AHEAD", "1") == "1":
                    mem.route_ahead(contrast_idx, self._feat_dim)
                v1, v2 = self._heads(f_s, f_t, reducer)
                return _ShardedPeerFn.apply(v1, v2, mem, idx, contrast_idx, self.criterion_s.n_data)
            except Exception as e:                  # no peer mapping on this system: fall back, loudly, once
                if mem._peer:
                    raise
                print(f"[mml_b200] symmetric-memory peer transport unavailable ({type(e).__name__}: {e}); using all_to_all")
                self.transport = "alltoall"
        handle = mem.exchange_begin(contrast_idx)                                   # routing kernels + size exchange
        v1, v2 = self._heads(f_s, f_t)                                              # local anchors
        D = v1.shape[1]
        V = _AllGatherRows.apply(torch.cat((v1, v2), 1), g)                         # [B_global, 2D]
        YP = _all_gather(torch.stack((idx, contrast_idx[:, 0]), 1), g)              # anchors' ids | positives' rows
        routed_ids, seg_ptr = mem.exchange_finish(handle)                           # host sync lands here
        routed = (routed_ids, seg_ptr, YP[:, 1].contiguous())
        return mem.fused_nce_loss(V[:, :D], V[:, D:], YP[:, 0].contiguous(), routed, self.criterion_s.n_data)
