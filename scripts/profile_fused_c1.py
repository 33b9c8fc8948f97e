#!/usr/bin/env python
"""Kernel timeline of ONE graph-replayed fused step at BASELINE config 1 (BilinearFusion 32x32->64 -> CRDLoss, batch 64):
which kernels the 0.37 ms consist of and where the stream idles (torch.profiler)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import multimodal_learning_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(2019)
B, dims, N, D, K, n = 64, (32, 32), 64, 128, 4096, 4096
fusion = pkg.BilinearFusion(skip=0, dim1=32, dim2=32, mmhid=64, dropout_rate=0.25).to(dev).train()
crd = pkg.CRDLoss(bench.make_opt(dict(B=B, D=D, K=K, n=n, s_dim=N, t_dim=N), n)).to(dev)
params = list(fusion.parameters()) + list(crd.parameters())
gen = torch.Generator(device=dev).manual_seed(7)
vecs = [torch.randn(B, d, device=dev, generator=gen) for d in dims]
f_t = torch.randn(B, N, device=dev, generator=gen)
idx = torch.randperm(n, device=dev, generator=gen)[:B].contiguous()
cidx = torch.randint(0, n, (B, K + 1), device=dev, generator=gen)
cidx[:, 0] = idx
optim = torch.optim.Adam(params, lr=2e-4, fused=True, capturable=True)
gstep = pkg.GraphedTrainStep(lambda a, b, c, d, e: crd(fusion(a, b), c, d, e), params, optim, (*vecs, f_t, idx, cidx), warmup=3, n_buffers=1)
for _ in range(10):
    gstep.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        gstep.replay()
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
per = len(ev) // 4
step = ev[2 * per:3 * per]
t0 = step[0].time_range.start
end = t0
busy = 0.0
print(f"# one replayed step = {per} kernels, span {(max(e.time_range.end for e in step) - t0):.1f} us")
for e in step:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    print(f"{s:8.1f} {d:7.1f} {e.time_range.start - end:7.1f}  {e.name[:100]}")
    end = max(end, e.time_range.end)
