#!/usr/bin/env python
"""Selection variant of CRD (5-arg CRDLoss over ContrastMemory_v3) at two sizes:
  ref   the reference's defaults (options.py:85-91,136): batch 16, P=300, K=700, P2=10, K2=512, n_data 1024
  big   config-2 scale: batch 1024, P=300, K=16384, P2=10, K2=8192, n_data 1M, feat_dim 128
Per size: ms of the relation-diff kernel (+ achieved GB/s over its algorithmic bytes 2*B*(K+P)*D*4 + B*(K+P)*12), of the
sort/top-k selection, of the fused multi-positive loss kernel, and of the whole step (forward + backward + Adam)."""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multimodal_learning_b200 as pkg
from multimodal_learning_b200 import crd_select as cs

dev = torch.device("cuda:0")


def ev_ms(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(name, B, P, K, P2, K2, n, D=128, iters=20):
    torch.manual_seed(0)
    np.random.seed(0)
    opt = types.SimpleNamespace(s_dim=128, t_dim=128, feat_dim=D, nce_p=P, nce_p2=P2, nce_k=K, nce_k2=K2, nce_t=0.07, nce_m=0.5,
                                select_pos_pairs=True, select_neg_pairs="True", sample_KD="False", select_pos_mode="hard")
    mod = cs.CRDLoss(opt, n).to(dev)
    params = list(mod.parameters())
    optim = torch.optim.Adam(params, lr=2e-4, fused=True)
    gen = torch.Generator(device=dev).manual_seed(1)
    pool = []
    for _ in range(3):
        idx = torch.randperm(n, device=dev, generator=gen)[:B].contiguous()
        cidx = torch.randint(0, n, (B, P + K), device=dev, generator=gen)
        cidx[:, 0] = idx
        pool.append((torch.randn(B, 128, device=dev, generator=gen), torch.randn(B, 128, device=dev, generator=gen), idx, cidx))
    it = [0]

    def step():
        f_s, f_t, idx, cidx = pool[it[0] % 3]
        it[0] += 1
        f_s = f_s.detach().requires_grad_(True)
        for p in params:
            p.grad = None
        loss = mod(0.0, f_s, f_t, idx, cidx)
        loss.backward()
        optim.step()
    step_ms = ev_ms(step, iters)
    mem = mod.contrast
    v1 = torch.nn.functional.normalize(pool[0][0], dim=1)
    v2 = torch.nn.functional.normalize(pool[0][1], dim=1)
    cidx = pool[0][3]
    rel_ms = ev_ms(lambda: cs.crd_relation_diff(mem.memory_v1, mem.memory_v2, v1, v2, cidx), iters)
    diff = cs.crd_relation_diff(mem.memory_v1, mem.memory_v2, v1, v2, cidx)
    sel_ms = ev_ms(lambda: (torch.sort(diff[:, :P], dim=1, descending=True), torch.topk(diff[:, P:], K2, dim=1, largest=False)), iters)
    own_ms = ev_ms(lambda: (cs.sort_columns(diff, 0, P, descending=True),
                            cs.sort_columns(diff, P, K, descending=False, first=K2, label0=P)), iters)
    _, sel_idx = mem.select(0.0, v1, v2, cidx, "hard")
    fused_ms = ev_ms(lambda: cs.crd_fused_loss_grad_multipos(mem.memory_v1, mem.memory_v2, v1, v2, sel_idx, P2, mem._T,
                                                            mem.params[2:4], n), iters)
    rel_bytes = 2 * B * (K + P) * D * 4 + B * (K + P) * 12
    fused_bytes = 2 * B * (P2 + K2) * D * 4 + B * (P2 + K2) * 8
    print(json.dumps({"config": name, "B": B, "P": P, "K": K, "P2": P2, "K2": K2, "n_data": n, "step_ms": round(step_ms, 4),
                      "steps_per_s": round(1000 / step_ms, 1), "relation_ms": round(rel_ms, 4),
                      "relation_GBps": round(rel_bytes / rel_ms / 1e6, 1), "library_sort_topk_ms": round(sel_ms, 4), "select_sort_kernel_ms": round(own_ms, 4),
                      "fused_multipos_ms": round(fused_ms, 4), "fused_GBps": round(fused_bytes / fused_ms / 1e6, 1)}), flush=True)


if os.environ.get("MML_SELECT_ONLY") != "big":        # ncu captures set this to profile the config-2-scale launches only
    run("ref defaults (batch 16, P300 K700 -> P2 10, K2 512, n 1024)", 16, 300, 700, 10, 512, 1024, iters=100)
run("config-2 scale (batch 1024, P300 K16384 -> P2 10, K2 8192, n 1M)", 1024, 300, 16384, 10, 8192, 1_000_000, iters=20)
