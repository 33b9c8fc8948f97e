"""Worker for the sharded-bank tests: runs the golden CRD steps on `world` ranks and checks every
rank's results against the reference golden (single-process, single-bank) values.

  backend "oracle": CPU tensors + gloo, compute steps by the CPU oracle  (host-logic test, no GPU)
  backend "cuda":   one GPU per rank + NCCL, compute steps by libmml_b200.so
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleBackend:
    """Stands in for the CUDA kernels on CPU so the collective choreography can run under gloo."""

    def route(self, cidx, rows_per, world):
        B, cols = cidx.shape
        owner = (cidx // rows_per).numpy()
        c = cidx.numpy()
        counts = np.zeros((world, B), dtype=np.int64)
        chunks = []
        for o in range(world):
            for b in range(B):
                sel = c[b][owner[b] == o]
                counts[o, b] = sel.size
                chunks.append((sel - o * rows_per).astype(np.int32))
        ids = np.concatenate(chunks) if chunks else np.zeros(0, np.int32)
        return torch.from_numpy(counts), torch.from_numpy(ids)

    @staticmethod
    def _segments(ids, seg_ptr):
        sp = seg_ptr.tolist()
        return [ids[sp[i]:sp[i + 1]].long() for i in range(len(sp) - 1)]

    def stats(self, bank1, bank2, v1, v2, ids, seg_ptr, T, cols):
        s = torch.zeros(4)
        for b, rows in enumerate(self._segments(ids, seg_ptr)):
            if rows.numel():
                s[2] += torch.exp(bank2[rows] @ v1[b] / T).sum()
                s[3] += torch.exp(bank1[rows] @ v2[b] / T).sum()
        return s

    def fused(self, bank1, bank2, v1, v2, ids, seg_ptr, pos_flag, T, Z, n_data, nce_k, batch):
        from oracle.crd_oracle import NCE_EPS
        kp = nce_k * (1 / float(n_data))
        c = kp + NCE_EPS
        sums = torch.zeros(4, dtype=torch.float64)
        g1 = torch.zeros(v1.shape, dtype=torch.float64)
        g2 = torch.zeros(v2.shape, dtype=torch.float64)
        for b, rows in enumerate(self._segments(ids, seg_ptr)):
            if not rows.numel():
                continue
            is_pos = torch.zeros(rows.numel(), dtype=torch.bool)
            is_pos[0] = bool(pos_flag[b])
            for side, (bank, v, z, g) in enumerate(((bank2, v1, Z[0], g1), (bank1, v2, Z[1], g2))):
                w = bank[rows].double()
                x = torch.exp(w @ v[b].double() / T) / float(z)
                term = torch.where(is_pos, torch.log(x / (x + c)), torch.log(kp / (x + c)))
                coef = torch.where(is_pos, -c / (x + c), x / (x + c)) / (T * batch)
                sums[side] += term.sum()
                g[b] = (coef.unsqueeze(1) * w).sum(0)
        return sums.float(), g1.float(), g2.float()

    def update(self, bank1, bank2, v1, v2, y, momentum, row_begin, row_end):
        from oracle.crd_oracle import momentum_update_
        own = ((y >= row_begin) & (y < row_end)).nonzero().flatten()
        if own.numel():
            momentum_update_(bank1, y[own] - row_begin, v1[own], momentum)
            momentum_update_(bank2, y[own] - row_begin, v2[own], momentum)


def run(rank, world, backend, golden_name, port, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if backend == "cuda":
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import multimodal_learning_b200.crd as _crd_mod
        _crd_mod._ALLOW_HOST_NORMALIZE = True      # host-logic test: the Embed heads run on CPU tensors here, nowhere else
    from conftest import Golden, rel_err
    from multimodal_learning_b200.sharded import ShardedCRDLoss
    g = Golden(golden_name)
    c = g.cfg
    assert c["B"] % world == 0
    Bl = c["B"] // world
    opt = types.SimpleNamespace(s_dim=c["s_dim"], t_dim=c["t_dim"], feat_dim=c["D"], n_data=c["n"], nce_k=c["K"],
                                nce_t=0.07, nce_m=0.5)
    mod = ShardedCRDLoss(opt, device=dev if backend == "cuda" else None,
                         backend=OracleBackend() if backend == "oracle" else None,
                         transport=os.environ.get("MML_TRANSPORT", "auto"))
    sd = g.state_dict("init.")
    mod.embed_s.load_state_dict({k[len("embed_s."):]: v for k, v in sd.items() if k.startswith("embed_s.")})
    mod.embed_t.load_state_dict({k[len("embed_t."):]: v for k, v in sd.items() if k.startswith("embed_t.")})
    mod.contrast.load_full_banks(sd["contrast.memory_v1"].to(dev), sd["contrast.memory_v2"].to(dev))
    tol = 1e-4
    worst = 0.0
    for s in range(c["steps"]):
        p = f"step{s}."
        sl = slice(rank * Bl, (rank + 1) * Bl)
        f_s = g.t(p + "f_s")[sl].to(dev).requires_grad_(True)
        f_t = g.t(p + "f_t")[sl].to(dev).requires_grad_(True)
        idx = g.t(p + "idx")[sl].to(dev)
        cidx = g.t(p + "contrast_idx")[sl].contiguous().to(dev)
        mod.zero_grad()
        loss = mod(f_s, f_t, idx, cidx)
        loss.backward()
        errs = {
            "loss": rel_err(loss, g.t(p + "loss")),
            "grad_f_s": rel_err(f_s.grad, g.t(p + "grad_f_s")[sl]),
            "grad_f_t": rel_err(f_t.grad, g.t(p + "grad_f_t")[sl]),
            "params": rel_err(mod.contrast.params, g.t(p + "params")),
        }
        for k, v in mod.named_parameters():
            errs["grad." + k] = rel_err(v.grad, g.t(p + "grad." + k))
        m1, m2 = mod.contrast.gather_full_banks()
        errs["memory_v1"] = rel_err(m1, g.t(p + "memory_v1"))
        errs["memory_v2"] = rel_err(m2, g.t(p + "memory_v2"))
        for k, e in errs.items():
            assert e < tol, f"rank {rank} step {s} {k}: rel {e:.3e}"
            worst = max(worst, e)
    dist.barrier()
    if rank == 0:
        with open(result_path, "w") as f:
            f.write(f"ok {worst:.3e} transport={mod.transport}\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    world, backend, name, port, out = int(sys.argv[1]), sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
    mp.spawn(run, args=(world, backend, name, port, out), nprocs=world, join=True)
