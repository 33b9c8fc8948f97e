#!/usr/bin/env python
"""Does the 135 MB H2D upload of the next step slow down the replayed step it overlaps?  (e2e = 2.84 ms vs value 2.59 ms.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import multimodal_learning_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
cfg = dict(bench.C2)
n = cfg["n"]
mod = pkg.CRDLoss(bench.make_opt(cfg, n)).to(dev)
params = list(mod.parameters())
optim = torch.optim.Adam(params, lr=2e-4, fused=True, capturable=True)
gen = torch.Generator(device=dev).manual_seed(0)
pool = [bench.gen_inputs(cfg, cfg["B"], n, gen, dev) for _ in range(2)]
gs = pkg.GraphedTrainStep(lambda a, b, c, d: mod(a, b, c, d), params, optim, pool[0], grad_inputs=(0,), warmup=3, n_buffers=2)
for slot, entry in enumerate(pool):
    for dst, src in zip(gs.buffers(slot), entry):
        dst.detach().copy_(src)
host = torch.empty(cfg["B"], cfg["K"] + 1, dtype=torch.int64).pin_memory()
host.random_(0, n)
scratch = torch.empty_like(pool[0][3])
copy_stream = torch.cuda.Stream(dev)


def timed(fn, it=20):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def with_copy():
    with torch.cuda.stream(copy_stream):
        scratch.copy_(host, non_blocking=True)
    gs.replay()


def with_copy_and_sync():
    with torch.cuda.stream(copy_stream):
        scratch.copy_(host, non_blocking=True)
    return gs.replay().item()


print("replay alone            %.3f ms" % timed(lambda: gs.replay()))
print("replay + item()         %.3f ms" % timed(lambda: gs.replay().item()))
print("replay under H2D        %.3f ms" % timed(with_copy))
print("replay under H2D + item %.3f ms" % timed(with_copy_and_sync))
print("H2D alone               %.3f ms" % timed(lambda: scratch.copy_(host, non_blocking=True)))
