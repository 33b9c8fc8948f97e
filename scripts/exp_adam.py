#!/usr/bin/env python
"""The fused Adam kernel is 69 of the 373 us of the replayed config-1 step: compare optimizer implementations on the same
parameter set (fused / foreach, capturable), replayed from a CUDA graph."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import multimodal_learning_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(2019)
fusion = pkg.BilinearFusion(skip=0, dim1=32, dim2=32, mmhid=64, dropout_rate=0.25).to(dev).train()
crd = pkg.CRDLoss(bench.make_opt(dict(B=64, D=128, K=4096, n=4096, s_dim=64, t_dim=64), 4096)).to(dev)
params = list(fusion.parameters()) + list(crd.parameters())
print("tensors", len(params), "elements", sum(p.numel() for p in params), "largest", max(p.numel() for p in params))
for p in params:
    p.grad = torch.randn_like(p)
for kw in (dict(fused=True), dict(foreach=True), dict(fused=True, capturable=False)):
    try:
        opt = torch.optim.Adam(params, lr=2e-4, capturable=kw.pop("capturable", True), **kw)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                opt.step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            opt.step()
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(kw, "%.1f us per replayed optimizer step" % (e0.elapsed_time(e1) / 50 * 1e3))
    except Exception as e:  # noqa: BLE001
        print(kw, "failed:", type(e).__name__, str(e)[:120])
