"""Worker for the sharded k-means test: `sharded_class_kmeans` over row shards on N GPUs must land on the centres of
`class_kmeans` over the whole bank on one GPU from the same start; its own k-means++ start must give every rank the same,
converged centres."""
from __future__ import annotations

import os
import sys

import numpy as np
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
    km = pkg.crd_kmeans
    worst = 0.0
    for n, D, k, lopsided in ((30000, 128, 3, False), (9001, 64, 5, True), (4000, 32, 2, False)):
        rng = np.random.default_rng(5)                                # the same bank on every rank
        modes = rng.standard_normal((9, D)).astype(np.float32)
        X = modes[rng.integers(0, 9, n)] + 0.3 * rng.standard_normal((n, D)).astype(np.float32)
        labels = rng.integers(0, 3, n)
        if lopsided:                                                  # class 2 lives on the first shard only
            labels[n // world:] = rng.integers(0, 2, n - n // world)
        class_idx = [np.nonzero(labels == c)[0] for c in range(3)]
        init = np.stack([X[rng.choice(r, k, replace=False)] for r in class_idx])
        bank = torch.from_numpy(X).to(dev)
        want, info = km.class_kmeans(bank, km.ClassRows(class_idx, dev), k, init=torch.from_numpy(init), return_info=True)
        assert bool(info["done"].all())
        per = (n + world - 1) // world
        lo, hi = rank * per, min(n, (rank + 1) * per)
        local_idx = [r[(r >= lo) & (r < hi)] - lo for r in class_idx]
        cls_local = km.ClassRows(local_idx, dev, allow_empty=True)
        shard = bank[lo:hi].contiguous()
        got, sinfo = km.sharded_class_kmeans(shard, cls_local, k, init=torch.from_numpy(init), return_info=True)
        assert bool(sinfo["done"].all())
        err = float((got - want).abs().max())
        assert err < 2e-5, f"rank {rank} n={n}: centres differ by {err:.3e}"
        assert float((sinfo["tol"] / info["tol"] - 1).abs().max()) < 1e-4
        worst = max(worst, err)
        # its own start: drawn by rank 0, rows supplied by their owners -> identical, converged centres on every rank
        gen = torch.Generator(device=dev)
        gen.manual_seed(3 + rank)                                     # only rank 0's draws count
        own, oinfo = km.sharded_class_kmeans(shard, cls_local, k, generator=gen, return_info=True)
        assert bool(oinfo["done"].all()) and bool(torch.isfinite(own).all())
        ref = own.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, own), f"rank {rank}: centres differ between ranks"
    dist.barrier()
    pkg.check_device_errors()
    if rank == 0:
        with open(result_path, "w") as f:
            f.write(f"ok {worst:.3e}\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    mp.spawn(run, args=(world, port, out), nprocs=world, join=True)
