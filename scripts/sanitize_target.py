#!/usr/bin/env python
"""compute-sanitizer target: every kernel family once, at small shapes.

    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_target.py [kron|crd|select|all]"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import multimodal_learning_b200 as pkg  # noqa: E402
from multimodal_learning_b200.fusion import KronLinearState, kron_linear  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
torch.manual_seed(0)
if what in ("kron", "all"):
    for B, dims, N, p in ((300, (32, 32), 64, 0.0), (200, (16, 24), 40, 0.25), (130, (8, 6, 10), 24, 0.25), (256, (32, 32, 32), 96, 0.0)):
        kk = 1
        for d in dims:
            kk *= d + 1
        fs = [torch.rand(B, d, device=dev).requires_grad_(True) for d in dims]
        W = (torch.randn(N, kk, device=dev) / kk ** 0.5).requires_grad_(True)
        st = KronLinearState(dims)
        y = kron_linear(st, fs, W, torch.zeros(N, device=dev), drop_p=p, training=p > 0, seed=3)
        y.sum().backward()
        torch.cuda.synchronize()
        print("kron", B, dims, N, p, "ok", float(y.abs().mean()), flush=True)
if what in ("crd", "all"):
    for D, K, n, B in ((128, 300, 2000, 16), (64, 100, 500, 8), (48, 64, 300, 6)):
        opt = types.SimpleNamespace(s_dim=32, t_dim=32, feat_dim=D, n_data=n, nce_k=K, nce_t=0.07, nce_m=0.5)
        crd = pkg.CRDLoss(opt).to(dev)
        for step in range(2):
            f_s = torch.randn(B, 32, device=dev, requires_grad=True)
            f_t = torch.randn(B, 32, device=dev)
            idx = torch.randperm(n, device=dev)[:B]
            cidx = torch.randint(0, n, (B, K + 1), device=dev)
            cidx[:, 0] = idx
            loss = crd(f_s, f_t, idx, cidx)
            loss.backward()
        loss2 = crd(f_s.detach(), f_t, idx, None)            # AliasMethod draw path
        o1, o2 = crd.contrast(torch.nn.functional.normalize(torch.randn(B, D, device=dev)),
                              torch.nn.functional.normalize(torch.randn(B, D, device=dev)), idx, cidx)
        torch.cuda.synchronize()
        print("crd", D, K, n, "ok", float(loss), float(loss2), flush=True)
    # selection variant + sampler
    labels = torch.randint(0, 3, (600,))
    samp = pkg.InstanceSampler(labels, nce_k=64).cuda(dev)
    rows = samp(torch.arange(8, device=dev))
    torch.cuda.synchronize()
    print("sampler ok", tuple(rows.shape), flush=True)
if what in ("select", "all"):
    # selection variants: relation gaps, per-anchor sort (stage-by-stage and register-blocked), multi-positive loss
    from multimodal_learning_b200 import _cabi, crd_knn, crd_select
    for n_cols in (300, 2500):
        diff = torch.randn(6, n_cols + 10, device=dev)
        order = crd_select.sort_columns(diff, 5, n_cols, descending=True)
        ref = torch.sort(diff[:, 5:5 + n_cols], dim=1, descending=True, stable=True)[1]
        print("sort", n_cols, "ok", bool(torch.equal(order, ref)), flush=True)
    opt = types.SimpleNamespace(s_dim=24, t_dim=24, feat_dim=64, nce_p=40, nce_p2=5, nce_k=200, nce_k2=64, nce_t=0.07, nce_m=0.5,
                                select_pos_pairs=True, select_neg_pairs="True", sample_KD="False", select_pos_mode="hard")
    sel = crd_select.CRDLoss(opt, 700).to(dev)
    for step in range(2):
        f_s = torch.randn(8, 24, device=dev, requires_grad=True)
        idx = torch.randperm(700, device=dev)[:8]
        cidx = torch.randint(0, 700, (8, 240), device=dev)
        cidx[:, 0] = idx
        sel(0.0, f_s, torch.randn(8, 24, device=dev), idx, cidx).backward()
    torch.cuda.synchronize()
    print("selection ok", flush=True)
    # full-bank KNN positives: tensor pass (one and two anchor tiles per CTA, with and without the class table), the
    # sampling pass (n >= 32768 rows), the exact scans
    for n, D, B, ncls in ((3000, 128, 20, 3), (40000, 64, 150, 3), (2500, 96, 9, 0), (1500, 48, 7, 0)):
        bank = torch.randn(n, D, device=dev)
        labels = torch.randint(0, 3, (n,), device=dev, dtype=torch.int32)
        rows = torch.randperm(n, device=dev)[:B]
        for exact in (False, True):
            nbr, sim, flags = crd_knn.knn_positives(bank, labels, rows, labels[rows].long(), 4, n_classes=ncls, exact_only=exact,
                                                    return_flags=True)
        torch.cuda.synchronize()
        print("knn", n, D, B, "ok", int(flags.sum()), bool((nbr[:, 0] == rows).all()), flush=True)
    # routing kernels (blocked and warp-per-slot) of the sharded bank, emulated world
    lib = _cabi.lib()
    for world, chunk in ((4, 2048), (3, 96)):
        Bq, cols, rows_per = 5, 4500, 1000
        cidx = torch.randint(0, rows_per * world, (Bq, cols), device=dev)
        chunks = (cols + chunk - 1) // chunk
        counts = torch.zeros(world * Bq * chunks, dtype=torch.int32, device=dev)
        ids = torch.zeros(world * Bq * chunks * chunk, dtype=torch.int32, device=dev)
        _cabi.check(lib.mml_shard_route_strided(_cabi.dptr(cidx, torch.int64), Bq, cols, chunk, rows_per, world, _cabi.dptr(counts),
                                                _cabi.dptr(ids), _cabi.cur_stream(dev)), "route")
        torch.cuda.synchronize()
        print("route", world, chunk, "ok", int(counts.sum()) == Bq * cols, flush=True)
if what in ("kmeans", "all"):
    # per-class k-means centres: every feature width (CTA sizes 256 / 128 / 64), ragged classes, the stop flags
    from multimodal_learning_b200 import crd_kmeans
    for D, k, sizes in ((32, 8, (300, 17, 1000)), (128, 3, (1000, 777, 1)), (128, 7, (2000, 900, 1500)), (64, 5, (700, 300, 90)), (256, 4, (900, 31, 650)), (512, 2, (260, 100, 7))):
        n = sum(sizes)
        bank = torch.randn(n, D, device=dev)
        perm = torch.randperm(n).numpy()
        class_idx, at = [], 0
        for m in sizes:
            class_idx.append(perm[at:at + m])
            at += m
        cls = crd_kmeans.ClassRows(class_idx, dev)
        centres, info = crd_kmeans.class_kmeans(bank, cls, k, max_iter=24, return_info=True)
        torch.cuda.synchronize()
        print("kmeans", D, k, "ok", bool(torch.isfinite(centres).all()), info["done"].tolist(), flush=True)
pkg.check_device_errors()
print("done", flush=True)
