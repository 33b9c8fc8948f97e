"""Worker for the sharded CUDA-graph test: the row-sharded step (routing, NVLink pulls, symmetric-memory barriers and the
NCCL-free head-gradient reduction) replayed from `GraphedTrainStep` must equal the same steps launched eagerly -- also
when the next batch's indices are routed one step early (`next_contrast_idx`), eagerly and from replayed graphs."""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(rank, world, port, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import multimodal_learning_b200 as pkg
    from conftest import rel_err
    from multimodal_learning_b200.sharded import ShardedCRDLoss
    Bl, D, K, n = 8, 128, 256, 1000 * world
    opt = types.SimpleNamespace(s_dim=40, t_dim=24, feat_dim=D, n_data=n, nce_k=K, nce_t=0.07, nce_m=0.5)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)

    def inputs():
        idx = torch.randperm(n, device=dev, generator=gen)[:Bl].contiguous()
        cidx = torch.randint(0, n, (Bl, K + 1), device=dev, generator=gen)
        cidx[:, 0] = idx
        return (torch.randn(Bl, 40, device=dev, generator=gen), torch.randn(Bl, 24, device=dev, generator=gen), idx, cidx)
    batches = [inputs() for _ in range(7)]
    results = []
    for mode in ("eager", "graph", "eager_prefetch", "graph_prefetch"):
        torch.manual_seed(5)                      # same heads and the same bank shards in both arms
        mod = ShardedCRDLoss(opt, device=dev)
        params = list(mod.parameters())
        # SGD: Adam's g/(|g|+eps) turns last-bit differences (cuBLAS may pick another algorithm under capture) into +-lr
        optim = torch.optim.SGD(params, lr=0.05, momentum=0.9)
        losses, grads = [], []
        if mode == "graph":
            step = pkg.GraphedTrainStep(lambda a, b, i, ci: mod(a, b, i, ci), params, optim, batches[0], grad_inputs=(0,),
                                        warmup=1, n_buffers=2)
            torch.cuda.synchronize()
            dist.barrier()
            for b in batches[1:]:
                losses.append(step(*b).clone())
                grads.append(step.static_in[(step._next - 1) % 2][0].grad.clone())
        elif mode == "graph_prefetch":
            # replay() over pre-filled buffers, round robin: graph k routes slot k+1's indices under its own tail, so slot
            # k+1 must hold its batch BEFORE graph k runs; slot 0's first batch is routed by hand (the capture baked in
            # "already routed")
            step = pkg.GraphedTrainStep(lambda a, b, i, ci, next_inputs=None: mod(a, b, i, ci, next_contrast_idx=next_inputs[3]),
                                        params, optim, batches[0], grad_inputs=(0,), warmup=1, n_buffers=2, pass_next_inputs=True)
            torch.cuda.synchronize()
            dist.barrier()

            def fill(slot, b):
                for dst, src in zip(step.buffers(slot), b):
                    dst.detach().copy_(src)
            todo = batches[1:]
            fill(0, todo[0])
            fill(1, todo[1])
            mod.contrast.prefetch_routing(step.buffers(0)[3], D)
            pkg.graphed.run_deferred()
            for i in range(len(todo)):
                slot = i % 2
                losses.append(step.replay(slot).clone())
                grads.append(step.static_in[slot][0].grad.clone())
                if i + 2 < len(todo):
                    fill(slot, todo[i + 2])
        else:
            for i, b in enumerate(batches):
                f_s = b[0].clone().requires_grad_(True)
                for p in params:
                    p.grad = None
                nxt = batches[i + 1][3] if (mode == "eager_prefetch" and i + 1 < len(batches)) else None
                loss = mod(f_s, *b[1:], next_contrast_idx=nxt) if mode == "eager_prefetch" else mod(f_s, *b[1:])
                loss.backward()
                optim.step()
                pkg.graphed.run_deferred()
                if i > 0:
                    losses.append(loss.detach().clone())
                    grads.append(f_s.grad.clone())
                del loss
        torch.cuda.synchronize()
        dist.barrier()
        m1, m2 = mod.contrast.gather_full_banks()
        results.append((losses, grads, m1.clone(), [p.detach().clone() for p in params]))
    (l0, g0, b0, p0), (l1, g1, b1, p1) = results[0], results[1]
    worst = 0.0
    for arm, (l1, g1, b1, p1) in zip(("graph", "eager_prefetch", "graph_prefetch"), results[1:]):
        named = ([(f"loss{i}", a, b) for i, (a, b) in enumerate(zip(l0, l1))]
                 + [(f"grad{i}", a, b) for i, (a, b) in enumerate(zip(g0, g1))]
                 + [("bank", b0, b1)] + [(f"param{i}", a, b) for i, (a, b) in enumerate(zip(p0, p1))])
        assert len(l1) == len(l0)
        for name, a, b in named:
            e = rel_err(b, a)
            assert e < 5e-5, f"rank {rank}: {arm} vs eager {name} rel {e:.3e}"
            worst = max(worst, e)
    p1 = results[-1][3]
    # replicated heads stay bit-identical across ranks (same operands, same order in the symmetric-memory reduction)
    w = p1[0].contiguous()
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    assert all(torch.equal(ws[0], x) for x in ws)
    dist.barrier()
    if rank == 0:
        with open(result_path, "w") as f:
            f.write(f"ok {worst:.3e}\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    mp.spawn(run, args=(world, port, out), nprocs=world, join=True)
