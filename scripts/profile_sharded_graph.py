#!/usr/bin/env python
"""Kernel timeline of ONE graph-replayed sharded CRD step (rank 0), under torchrun: start offset, duration and gap of every
kernel, from torch.profiler (CUPTI sees the kernels of a replayed graph).  Names the fixed per-step cost around the gather.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/profile_sharded_graph.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import multimodal_learning_b200 as pkg  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    cfg = dict(bench.C2)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from multimodal_learning_b200.sharded import ShardedCRDLoss
        n = bench.ROWS_PER_GPU_SHARDED * world
        mod = ShardedCRDLoss(bench.make_opt(cfg, n), device=dev)
    else:
        n = cfg["n"]
        mod = pkg.CRDLoss(bench.make_opt(cfg, n)).to(dev)
    params = list(mod.parameters())
    optim = torch.optim.Adam(params, lr=2e-4, fused=True, capturable=True)
    gen = torch.Generator(device=dev).manual_seed(rank)
    pool = [bench.gen_inputs(cfg, cfg["B"], n, gen, dev) for _ in range(2)]
    if world > 1 and os.environ.get("MML_PREFETCH_ROUTING", "1") == "1":      # as bench.py: next step's indices routed early
        gstep = pkg.GraphedTrainStep(lambda a, b, c, d, next_inputs=None: mod(a, b, c, d, next_contrast_idx=next_inputs[3]),
                                     params, optim, pool[0], grad_inputs=(0,), warmup=3, n_buffers=2, pass_next_inputs=True)
    else:
        gstep = pkg.GraphedTrainStep(lambda a, b, c, d: mod(a, b, c, d), params, optim, pool[0], grad_inputs=(0,), warmup=3,
                                     n_buffers=2)
    for slot, entry in enumerate(pool):
        for dst, src in zip(gstep.buffers(slot), entry):
            dst.detach().copy_(src)
    for _ in range(6):
        gstep.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            gstep.replay()
        torch.cuda.synchronize()
    if rank == 0:
        ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                    key=lambda e: e.time_range.start)
        gather = [i for i, e in enumerate(ev) if "crd_gather_kernel" in e.name]
        # the third replay: from the end of the second gather's step to the end of the third's
        per = len(ev) // 4
        step = ev[2 * per:3 * per]
        t0 = step[0].time_range.start
        prev_end = t0
        print(f"# world {world}: one replayed step = {per} kernels, span {(step[-1].time_range.end - t0):.1f} us")
        print(f"{'start_us':>9} {'dur_us':>8} {'gap_us':>7}  kernel")
        for e in step:
            s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
            print(f"{s:9.1f} {d:8.1f} {e.time_range.start - prev_end:7.1f}  {e.name[:110]}")
            prev_end = max(prev_end, e.time_range.end)
    if world > 1:
        dist.destroy_process_group()


main()
