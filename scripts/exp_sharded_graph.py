#!/usr/bin/env python
"""Does the sharded peer step survive CUDA-graph capture + replay?  Small sizes, stage prints (run under torchrun)."""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist

def log(*a):
    print(f"[r{os.environ.get('RANK')}] {time.strftime('%H:%M:%S')}", *a, file=sys.stderr, flush=True)

def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    import multimodal_learning_b200 as pkg
    from multimodal_learning_b200.sharded import ShardedCRDLoss
    B, D, K, n = 64, 128, 2048, 100_000 * world
    opt = types.SimpleNamespace(s_dim=128, t_dim=128, feat_dim=D, n_data=n, nce_k=K, nce_t=0.07, nce_m=0.5)
    torch.manual_seed(0)
    mod = ShardedCRDLoss(opt, device=dev)
    params = list(mod.parameters())
    optim = torch.optim.Adam(params, lr=1e-3, fused=True, capturable=True)
    gen = torch.Generator(device=dev).manual_seed(10 + rank)
    def inputs():
        idx = torch.randperm(n, device=dev, generator=gen)[:B].contiguous()
        cidx = torch.randint(0, n, (B, K + 1), device=dev, generator=gen); cidx[:, 0] = idx
        return (torch.randn(B, 128, device=dev, generator=gen), torch.randn(B, 128, device=dev, generator=gen), idx, cidx)
    pool = [inputs() for _ in range(4)]
    log("heads_reduce =", mod.heads_reduce)
    for i in range(3):
        for p in params: p.grad = None
        f_s = pool[i][0].clone().requires_grad_(True)
        loss = mod(f_s, *pool[i][1:]); loss.backward(); optim.step()
    torch.cuda.synchronize(); dist.barrier(); log("eager ok", loss.item())
    del loss, f_s
    gs = pkg.GraphedTrainStep(lambda a, b, c, d: mod(a, b, c, d), params, optim, pool[0], grad_inputs=(0,), warmup=2, n_buffers=2,
                              before_capture=lambda: log("capturing"))
    torch.cuda.synchronize(); log("captured"); dist.barrier(); torch.cuda.synchronize(); log("post-capture barrier ok")
    losses = []
    for i in range(20):
        losses.append(gs(*pool[i % 4]).clone())
    torch.cuda.synchronize(); log("replays ok", [round(l.item(), 4) for l in losses[:4]], losses[-1].item())
    w = params[0].detach().clone(); ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    log("params identical across ranks:", all(torch.equal(ws[0], x) for x in ws))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize(); e0.record()
    for i in range(50): gs.replay()
    e1.record(); torch.cuda.synchronize(); log("graph ms/step", e0.elapsed_time(e1) / 50)
    dist.destroy_process_group()
main()
