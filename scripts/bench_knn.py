#!/usr/bin/env python
"""Full-bank KNN positives (csrc/crd_knn.cu, reference `MIA 2023/.../CRD_criterion_v10.py:69-80`) at BASELINE config 2's bank:
n = 1M rows x 128, 1024 anchors, 5 positives.  CUDA-event time of the whole call (inverse norms + query prep + TF32 pass +
exact re-score + exact scan of flagged anchors), flops = 2 B n D against the TF32 peak, bytes = one pass over the bank.
Beside it: the reference's way on the host cores (sklearn cosine_similarity + sort) on a slice of anchors.

    python scripts/bench_knn.py [--cpu-anchors 8]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_learning_b200 import crd_knn  # noqa: E402


def run(n, D, B, P, iters, cpu_anchors, clustered=False):
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    bank = torch.randn(n, D, device=dev, generator=gen)
    if clustered:
        centres = torch.randn(64, D, device=dev, generator=gen)
        bank = centres[torch.randint(0, 64, (n,), device=dev, generator=gen)] + 0.05 * bank
    bank = bank / bank.norm(dim=1, keepdim=True)
    labels = torch.randint(0, 3, (n,), device=dev, generator=gen, dtype=torch.int32)
    rows = torch.randperm(n, device=dev, generator=gen)[:B]
    blab = labels[rows].long()
    for _ in range(3):
        idx, sim, flags = crd_knn.knn_positives(bank, labels, rows, blab, P, n_classes=3, return_flags=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        crd_knn.knn_positives(bank, labels, rows, blab, P, n_classes=3)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    inv = crd_knn.knn_inv_norms(bank)
    e0.record()
    for _ in range(iters):
        crd_knn.knn_positives(bank, labels, rows, blab, P, n_classes=3, inv_norms=inv)
    e1.record()
    torch.cuda.synchronize()
    ms_cached = e0.elapsed_time(e1) / iters
    out = {"workload": f"KNN positives: bank {n} x {D}, {B} anchors, num_pos {P}" + (", clustered rows" if clustered else ""),
           "ms": round(ms, 4), "ms_with_cached_norms": round(ms_cached, 4), "tflops_tf32": round(2.0 * B * n * D / ms / 1e9, 1),
           "bank_GBps": round(n * D * 4 * 2 / ms / 1e6, 1), "flagged_anchors": int(flags.sum())}
    if cpu_anchors > 0:
        from sklearn.metrics.pairwise import cosine_similarity
        hb, hl = bank.cpu(), labels.cpu().long()
        hr, hlab = rows[:cpu_anchors].cpu(), blab[:cpu_anchors].cpu()
        t0 = time.perf_counter()
        mask = (hl.view(1, -1) == hlab.view(-1, 1)).float()
        s = mask * torch.tensor(cosine_similarity(hb[hr].numpy(), hb.numpy()))
        order = torch.sort(s, descending=True, dim=-1)
        dt = time.perf_counter() - t0
        assert torch.equal(order[1][:, :P].to(dev), idx[:cpu_anchors]) or clustered
        out["cpu_reference"] = {"ms_per_call_scaled": round(dt * 1e3 * B / cpu_anchors, 1), "threads": torch.get_num_threads(),
                                "sample": f"{cpu_anchors} of {B} anchors (sklearn cosine_similarity + torch.sort over {n} columns), scaled x{B // cpu_anchors}"}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-anchors", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    run(1_000_000, 128, 1024, 5, a.iters, a.cpu_anchors)
    run(1_000_000, 128, 1024, 5, a.iters, 0, clustered=True)
    run(1024, 128, 16, 5, 100, 16)
