#!/usr/bin/env python
"""compute-sanitizer target: every kernel family once, at small shapes.

    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_target.py [kron|crd|all]"""
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
pkg.check_device_errors()
print("done", flush=True)
