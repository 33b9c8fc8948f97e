#!/usr/bin/env python
"""Fused distillation step (fusion -> CRD, forward + backward + bank update + Adam) at BASELINE configs 1 and 4.

  c1  BilinearFusion(path 32-d, omic 32-d, mmhid 64) teacher/student features, feat_dim 128, batch 64, n_data 4096,
      nce_k 4096 -- the reference's own CPU-runnable case; the CPU oracle port is timed beside it (host cores).
  c4  TrilinearFusion_A(32, 32, 32 -> 96), batch 8192, fused with the CRD loss (feat_dim 128, n_data 1M, nce_k 16384).

One JSON line per config.  Step = f = fusion(vecs); loss = CRDLoss(opt)(f, f_t, idx, contrast_idx); loss.backward();
Adam step over fusion + Embed parameters.  CUDA events, inputs resident on the device, fresh
idx/contrast_idx per step from a rotating pool.
"""
import argparse
import json
import os
import statistics
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import multimodal_learning_b200 as pkg  # noqa: E402


def run_gpu(name, make_fusion, dims, B, N, D, K, n, steps, warmup, graph=False):
    dev = torch.device("cuda:0")
    torch.manual_seed(2019)
    fusion = make_fusion().to(dev).train()
    opt = types.SimpleNamespace(s_dim=N, t_dim=N, feat_dim=D, n_data=n, nce_k=K, nce_t=0.07, nce_m=0.5)
    crd = pkg.CRDLoss(opt).to(dev)
    params = list(fusion.parameters()) + list(crd.parameters())
    optim = torch.optim.Adam(params, lr=2e-4, betas=(0.9, 0.999), fused=True)
    gen = torch.Generator(device=dev).manual_seed(7)
    pool = []
    for _ in range(4):
        vecs = [torch.randn(B, d, device=dev, generator=gen) for d in dims]
        f_t = torch.randn(B, N, device=dev, generator=gen)
        idx = torch.randperm(n, device=dev, generator=gen)[:B].contiguous()
        cidx = torch.randint(0, n, (B, K + 1), device=dev, generator=gen)
        cidx[:, 0] = idx
        pool.append((vecs, f_t, idx, cidx))

    def step(i):
        vecs, f_t, idx, cidx = pool[i % len(pool)]
        for p in params:
            p.grad = None
        loss = crd(fusion(*vecs), f_t, idx, cidx)
        loss.backward()
        optim.step()
        return loss

    if graph:       # whole step as ONE CUDA-graph launch; one captured graph per pool entry, no per-step input copies
        for i in range(3):
            step(i)
        optim = torch.optim.Adam(params, lr=2e-4, betas=(0.9, 0.999), fused=True, capturable=True)
        nv = len(dims)
        gstep = pkg.GraphedTrainStep(lambda *a: crd(fusion(*a[:nv]), *a[nv:]), params, optim,
                                     (*pool[0][0], *pool[0][1:]), warmup=2, n_buffers=len(pool))
        for slot, (vecs, f_t, idx, cidx) in enumerate(pool):
            for dst, src in zip(gstep.buffers(slot), (*vecs, f_t, idx, cidx)):
                dst.detach().copy_(src)

        def step(i):  # noqa: F811
            return gstep.replay(i % len(pool))

    for i in range(max(warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    l0 = pkg._cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kk = 1
    for d in dims:
        kk *= d + 1
    return {"config": name + (" [CUDA graph]" if graph else ""), "steps_per_s": 1000.0 / ms, "ms_per_step": ms, "B": B, "N": N, "K": K, "n_data": n,
            "kron_flops_fwd_bwd": 6 * B * kk * N, "crd_bytes": 2 * B * (K + 1) * D * 4 + B * (K + 1) * 8,
            "gpu_launches_per_step": (pkg._cabi.launch_count() - l0) / steps if not graph else "1 graph replay"}


def run_cpu_c1(steps, warmup):
    """The reference's op sequence (oracle port: fusion_oracle + crd_oracle) on the host cores, config 1."""
    from oracle import crd_oracle as co
    from oracle import fusion_oracle as fo
    torch.manual_seed(2019)
    B, d, N, D, K, n = 64, 32, 64, 128, 4096, 4096
    fusion = pkg.BilinearFusion(skip=0, dim1=d, dim2=d, mmhid=N, dropout_rate=0.0)
    sd_f = {k: v.clone() for k, v in fusion.state_dict().items()}
    opt = types.SimpleNamespace(s_dim=N, t_dim=N, feat_dim=D, n_data=n, nce_k=K, nce_t=0.07, nce_m=0.5)
    sd_c = {k: v.clone() for k, v in pkg.CRDLoss(opt).state_dict().items()}
    leaves = []
    for sd in (sd_f, sd_c):
        for k in sd:
            if sd[k].is_floating_point() and "running" not in k and "memory" not in k and "params" not in k:
                sd[k].requires_grad_(True)
                leaves.append(sd[k])
    optim = torch.optim.Adam(leaves, lr=2e-4)
    times = []
    for i in range(warmup + steps):
        v1, v2, f_t = torch.randn(B, d), torch.randn(B, d), torch.randn(B, N)
        idx = torch.randperm(n)[:B]
        cidx = torch.randint(0, n, (B, K + 1))
        cidx[:, 0] = idx
        t0 = time.perf_counter()
        f_s = fo.bilinear_fusion_forward(sd_f, v1, v2, skip=0, training=True)
        loss, _, _ = co.crd_loss(sd_c, f_s, f_t, idx, cidx, n)
        loss.backward()
        optim.step()
        optim.zero_grad(set_to_none=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return {"config": "c1_cpu_port", "steps_per_s": 1.0 / statistics.mean(times), "ms_per_step": 1e3 * statistics.mean(times),
            "cores": torch.get_num_threads()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run_gpu("c1 BilinearFusion(32,32->64) + CRD B=64 K=4096 n=4096",
                             lambda: pkg.BilinearFusion(skip=0, dim1=32, dim2=32, mmhid=64, dropout_rate=0.25),
                             (32, 32), 64, 64, 128, 4096, 4096, a.steps * 4, a.warmup)), flush=True)
    print(json.dumps(run_gpu("c1 BilinearFusion(32,32->64) + CRD B=64 K=4096 n=4096",
                             lambda: pkg.BilinearFusion(skip=0, dim1=32, dim2=32, mmhid=64, dropout_rate=0.25),
                             (32, 32), 64, 64, 128, 4096, 4096, a.steps * 20, a.warmup, graph=True)), flush=True)
    print(json.dumps(run_gpu("c4 TrilinearFusion_A(32^3->96) + CRD B=8192 K=16384 n=1M",
                             lambda: pkg.TrilinearFusion_A(skip=1, dim1=32, dim2=32, dim3=32, mmhid=96),
                             (32, 32, 32), 8192, 96, 128, 16384, 1_000_000, max(a.steps // 5, 5), 3)), flush=True)
    if not a.no_cpu:
        print(json.dumps(run_cpu_c1(20, 3)), flush=True)


if __name__ == "__main__":
    main()
