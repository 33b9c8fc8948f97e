"""Worker for the sharded KNN test: `sharded_knn_positives` over row shards on N GPUs must equal `knn_positives` over the whole
bank on one GPU (neighbour rows and similarities)."""
from __future__ import annotations

import os
import sys

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
    worst = 0.0
    for n, D, Bl, P, ncls in ((40000, 128, 24, 5, 3), (5001, 64, 8, 3, 0), (130, 32, 4, 8, 3)):
        gen = torch.Generator().manual_seed(11)                      # the same bank on every rank
        bank = (torch.randn(n, D, generator=gen) * (0.5 + torch.rand(n, 1, generator=gen))).to(dev)
        labels = torch.randint(0, 3, (n,), generator=gen).to(dev)
        rows_all = torch.randperm(n, generator=gen)[:world * Bl].to(dev)
        rows = rows_all[rank * Bl:(rank + 1) * Bl].contiguous()
        labs = labels[rows].long()
        want_idx, want_sim = pkg.crd_knn.knn_positives(bank, labels.int(), rows, labs, P, n_classes=ncls)
        per = (n + world - 1) // world
        lo, hi = rank * per, min(n, (rank + 1) * per)
        got_idx, got_sim = pkg.crd_knn.sharded_knn_positives(bank[lo:hi].contiguous(), lo, labels[lo:hi].int().contiguous(), rows, labs, P,
                                                             n_classes=ncls)
        assert torch.equal(got_idx, want_idx), f"rank {rank} n={n}: neighbours differ"
        err = float((got_sim - want_sim).abs().max())
        assert err < 1e-6, f"rank {rank} n={n}: similarity {err:.3e}"
        worst = max(worst, err)
    dist.barrier()
    if rank == 0:
        with open(result_path, "w") as f:
            f.write(f"ok {worst:.3e}\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    mp.spawn(run, args=(world, port, out), nprocs=world, join=True)
