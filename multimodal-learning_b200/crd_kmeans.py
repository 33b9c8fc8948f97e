"""Per-class k-means centres of a memory bank, on the device (csrc/crd_kmeans.cu, K13).

The reference's pos_extra == "centers" with num_pos > 2 (`MIA 2023/stage2_unimodal_student/CL_utils/CRD_criterion_v10.py:84-92`
for bank 1, `:122-129` for bank 2) copies every class's bank rows to the host and fits
`sklearn.cluster.KMeans(n_clusters=num_pos - 1)` on them -- in every forward, with sklearn's randomly seeded k-means++
initialisation, so the reference's own centres differ from run to run.  `class_kmeans` is the same estimator with its defaults
(k-means++ initialisation by D^2 sampling, Lloyd iterations, `max_iter=300`, `tol=1e-4` scaled by the mean feature variance,
one initialisation) as passes over the bank in HBM; given the same initial centres it reproduces sklearn's centres (the oracle
and `tests/golden/crdkmeans_*` pin that), and its random draws come from a torch generator, so it is reproducible under a seed.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi


class ClassRows:
    """The bank rows of every class, class after class, on the bank's device (+ the class boundaries on the host)."""

    def __init__(self, class_idx, device, allow_empty=False):
        lists = [torch.as_tensor(np.asarray(c), dtype=torch.long).reshape(-1) for c in class_idx]
        if not lists or (not allow_empty and any(r.numel() == 0 for r in lists)) or sum(r.numel() for r in lists) == 0:
            raise RuntimeError("every class needs at least one row")
        self.sizes = [int(r.numel()) for r in lists]
        self.offsets = torch.tensor([0] + list(np.cumsum(self.sizes)), dtype=torch.int64)      # host
        self.rows = torch.cat(lists).to(device)
        self.n_classes = len(lists)
        self.device = self.rows.device


def lloyd(bank, cls: ClassRows, centres, *, iterations=1, update=True, tol=None, done=None, inertia=None, counts=None,
          sums=None, row_dist=None, workspace=None):
    """`mml_crd_kmeans_lloyd`: `iterations` Lloyd iterations over the listed rows, centres [C, k, D] updated in place."""
    if not bank.is_cuda:
        raise RuntimeError("class k-means runs on CUDA tensors only")
    n, D = bank.shape
    C, k, Dc = centres.shape
    if C != cls.n_classes or Dc != D:
        raise RuntimeError(f"centres must be [{cls.n_classes}, k, {D}]; got {tuple(centres.shape)}")
    lib = _cabi.lib()
    if k > lib.mml_crd_kmeans_max_clusters():
        raise NotImplementedError(f"at most {lib.mml_crd_kmeans_max_clusters()} centres per class (got {k})")
    nbytes = lib.mml_crd_kmeans_workspace_bytes(C, k, D)
    if nbytes <= 0:
        raise NotImplementedError(f"class k-means: unsupported sizes classes={C} k={k} D={D}")
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=bank.device)
    _cabi.check(lib.mml_crd_kmeans_lloyd(
        _cabi.dptr(bank, torch.float32), n, D, _cabi.dptr(cls.rows, torch.int64), _cabi.hptr(cls.offsets), C, k,
        _cabi.dptr(centres, torch.float32), _cabi.dptr(tol, torch.float32), int(iterations), int(bool(update)),
        _cabi.dptr(done, torch.int32), _cabi.dptr(inertia, torch.float32), _cabi.dptr(counts, torch.int64),
        _cabi.dptr(sums, torch.float32), _cabi.dptr(row_dist, torch.float32), _cabi.dptr(workspace), workspace.numel(), _cabi.cur_stream(bank.device)),
        "mml_crd_kmeans_lloyd")
    return workspace


def class_variance_tolerance(bank, cls: ClassRows, tol=1e-4, workspace=None):
    """sklearn's `_tolerance` per class: tol * mean over features of the variance of the class's rows -> fp32 [C]."""
    C, D = cls.n_classes, bank.shape[1]
    mean = torch.zeros((C, 1, D), dtype=torch.float32, device=bank.device)
    inertia = torch.empty((C, 1), dtype=torch.float32, device=bank.device)
    ws = lloyd(bank, cls, mean, update=True, workspace=workspace)                 # one centre at 0 -> the class mean
    lloyd(bank, cls, mean, update=False, inertia=inertia, workspace=ws)           # sum |x - mean|^2
    sizes = torch.tensor(cls.sizes, dtype=torch.float32, device=bank.device)
    return (inertia[:, 0] / (sizes * D) * tol).contiguous()


def kmeans_plus_plus(bank, cls: ClassRows, k, generator=None, workspace=None):
    """k-means++ initial centres [C, k, D]: the first centre of a class is one of its rows drawn uniformly, every further
    one a row drawn with probability proportional to its squared distance to the nearest centre chosen so far."""
    dev = bank.device
    C, D = cls.n_classes, bank.shape[1]
    first = [cls.rows[int(cls.offsets[c]) + torch.randint(cls.sizes[c], (1,), generator=generator, device=dev)] for c in range(C)]
    centres = bank.index_select(0, torch.cat(first)).view(C, 1, D).repeat(1, k, 1).contiguous()   # duplicates do not change a min
    if k == 1:
        return centres
    row_dist = torch.empty(cls.rows.numel(), dtype=torch.float32, device=dev)
    for j in range(1, k):
        workspace = lloyd(bank, cls, centres, update=False, row_dist=row_dist, workspace=workspace)
        for c in range(C):
            lo, hi = int(cls.offsets[c]), int(cls.offsets[c + 1])
            cum = row_dist[lo:hi].double().cumsum(0)
            u = torch.rand(1, generator=generator, device=dev, dtype=torch.float64) * cum[-1]
            pick = torch.searchsorted(cum, u, right=True).clamp_(max=hi - lo - 1)
            centres[c, j] = bank[cls.rows[lo + pick]].view(D)
    return centres


def class_kmeans(bank, cls: ClassRows, k, *, init=None, generator=None, max_iter=300, tol=1e-4, check_every=8, return_info=False):
    """-> centres fp32 [C, k, D] (`KMeans(n_clusters=k).fit(rows of class c).cluster_centers_` for every class c).
    init: optional initial centres [C, k, D] (else k-means++ from `generator`).  The `done` flags are read back every
    `check_every` iterations -- the only host synchronisation."""
    bank = bank.detach()
    dev = bank.device
    C = cls.n_classes
    if init is not None and tuple(init.shape) != (C, k, bank.shape[1]):
        raise RuntimeError(f"init must be [{C}, {k}, {bank.shape[1]}]; got {tuple(init.shape)}")
    with torch.no_grad():
        tol_c = class_variance_tolerance(bank, cls, tol)
        centres = kmeans_plus_plus(bank, cls, k, generator) if init is None else init.to(device=dev, dtype=torch.float32).clone().contiguous()
        done = torch.zeros(C, dtype=torch.int32, device=dev)
        ws = None
        iters = 0
        while iters < max_iter:
            step = min(check_every, max_iter - iters)
            ws = lloyd(bank, cls, centres, iterations=step, tol=tol_c, done=done, workspace=ws)
            iters += step
            if bool(done.all()):
                break
    if return_info:
        return centres, {"done": done, "tol": tol_c, "iterations_enqueued": iters}
    return centres


def sharded_class_kmeans(bank_local, cls_local: ClassRows, k, *, group=None, init=None, generator=None, max_iter=300, tol=1e-4,
                         check_every=8, return_info=False):
    """`class_kmeans` over a bank whose rows are sharded across the ranks of `group`: `bank_local` is this rank's shard and
    `cls_local` lists ITS rows by class with local row numbers (`ClassRows(..., allow_empty=True)`: a class may be absent from
    a shard).  Bank rows never cross the wire: every Lloyd iteration each rank runs the assign pass over its shard, the ranks
    all-reduce the per-centre sums and counts ([C, k, D + 1] doubles) and apply the same update and stop rule; the k-means++
    draws are made by rank 0, the rank that owns the drawn row supplies it.  Returns the same centres on every rank."""
    import torch.distributed as dist
    bank = bank_local.detach()
    dev = bank.device
    C, D = cls_local.n_classes, bank.shape[1]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    cls = cls_local
    with torch.no_grad():
        sizes = torch.tensor(cls.sizes, dtype=torch.float64, device=dev)
        dist.all_reduce(sizes, group=group)
        if bool((sizes < 1).any()):
            raise RuntimeError("every class needs at least one row on some rank")
        # sklearn's tolerance: class mean, then the summed squared distance to it, both reduced over the shards
        mean = torch.zeros((C, 1, D), dtype=torch.float32, device=dev)
        sums1 = torch.empty((C, 1, D), dtype=torch.float32, device=dev)
        ws = lloyd(bank, cls, mean, update=False, sums=sums1)
        tot = sums1.double()
        dist.all_reduce(tot, group=group)
        mean = (tot / sizes.view(C, 1, 1)).float().contiguous()
        inertia = torch.empty((C, 1), dtype=torch.float32, device=dev)
        ws = lloyd(bank, cls, mean, update=False, inertia=inertia, workspace=ws)
        spread = inertia[:, 0].double()
        dist.all_reduce(spread, group=group)
        tol_c = (spread / (sizes * D) * tol).float().contiguous()

        def draw(weights):
            """One row per class, drawn with probability proportional to `weights` (per listed local row) over ALL shards."""
            cums = [weights[int(cls.offsets[c]):int(cls.offsets[c + 1])].double().cumsum(0) for c in range(C)]
            local = torch.stack([cm[-1] if cm.numel() else torch.zeros((), dtype=torch.float64, device=dev) for cm in cums])
            totals = torch.empty((world, C), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(totals, local.contiguous(), group=group)
            u = torch.rand(C, generator=generator, device=dev, dtype=torch.float64)
            dist.broadcast(u, src=src, group=group)
            upto = totals.cumsum(0)                                       # [world, C]
            target = u * upto[-1]
            rows = torch.zeros((C, D), dtype=torch.float32, device=dev)
            for c in range(C):
                owner = torch.searchsorted(upto[:, c].contiguous(), target[c:c + 1], right=True).clamp_(max=world - 1)
                if cums[c].numel() == 0:
                    continue
                before = upto[rank - 1, c] if rank > 0 else torch.zeros((), dtype=torch.float64, device=dev)
                pick = torch.searchsorted(cums[c], (target[c:c + 1] - before).clamp_(min=0), right=True).clamp_(max=cums[c].numel() - 1)
                row = bank[cls.rows[int(cls.offsets[c]) + pick]].view(D)
                rows[c] = torch.where(owner == rank, row, torch.zeros_like(row))
            dist.all_reduce(rows, group=group)
            return rows

        if init is None:
            n_listed = cls.rows.numel()
            centres = draw(torch.ones(n_listed, dtype=torch.float32, device=dev)).view(C, 1, D).repeat(1, k, 1).contiguous()
            row_dist = torch.empty(n_listed, dtype=torch.float32, device=dev)
            for j in range(1, k):
                ws = lloyd(bank, cls, centres, update=False, row_dist=row_dist, workspace=ws)
                centres[:, j] = draw(row_dist)
        else:
            if tuple(init.shape) != (C, k, D):
                raise RuntimeError(f"init must be [{C}, {k}, {D}]; got {tuple(init.shape)}")
            centres = init.to(device=dev, dtype=torch.float32).clone().contiguous()

        done = torch.zeros(C, dtype=torch.int32, device=dev)
        sums = torch.zeros((C, k, D), dtype=torch.float32, device=dev)
        counts = torch.zeros((C, k), dtype=torch.int64, device=dev)
        iters = 0
        while iters < max_iter:
            ws = lloyd(bank, cls, centres, update=False, done=done, sums=sums, counts=counts, workspace=ws)
            buf = torch.cat((sums.double(), counts.double().unsqueeze(2)), 2).contiguous()
            dist.all_reduce(buf, group=group)
            n_j = buf[:, :, D:]
            new = torch.where(n_j > 0, buf[:, :, :D] / n_j.clamp(min=1), centres.double()).float()      # an empty cluster keeps its centre
            shift = ((new - centres) ** 2).sum((1, 2))
            frozen = done.bool()
            centres = torch.where(frozen.view(C, 1, 1), centres, new).contiguous()
            done = (frozen | (shift <= tol_c)).int()
            iters += 1
            if iters % check_every == 0 and bool(done.all()):
                break
    if return_info:
        return centres, {"done": done, "tol": tol_c, "iterations_enqueued": iters}
    return centres
