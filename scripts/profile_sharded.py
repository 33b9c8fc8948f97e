#!/usr/bin/env python
"""torch.profiler view of the sharded CRD step (run under torchrun): per-kernel CUDA time on rank 0."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import bench

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
        import multimodal_learning_b200 as pkg
        n = cfg["n"]
        mod = pkg.CRDLoss(bench.make_opt(cfg, n)).to(dev)
    params = list(mod.parameters())
    optim = torch.optim.Adam(params, lr=2e-4, fused=True)
    gen = torch.Generator(device=dev).manual_seed(rank)
    pool = [bench.gen_inputs(cfg, cfg["B"], n, gen, dev) for _ in range(2)]
    def step(i):
        f_s, f_t, idx, cidx = pool[i % 2]
        f_s = f_s.detach().requires_grad_(True)
        for p in params: p.grad = None
        loss = mod(f_s, f_t, idx, cidx); loss.backward(); optim.step()
    for i in range(5): step(i)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    steps = 10
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(steps): step(i)
        torch.cuda.synchronize()
    if rank == 0:
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
        busy = sum(e.time_range.end - e.time_range.start for e in ev)
        print(f"span {(t1-t0)/steps:.1f} us/step, sum of kernel time {busy/steps:.1f} us/step")
    if world > 1: dist.destroy_process_group()
main()
