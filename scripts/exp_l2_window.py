#!/usr/bin/env python
"""Experiment: K4 throughput when all gathered rows fall in an L2-sized window of the banks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_learning_b200 import crd
dev = torch.device("cuda:0")
B, D, K = 1024, 128, 16384
T = 0.07
Z = torch.tensor([1e6, 1e6], device=dev)
for n in (16384, 32768, 65536, 131072, 262144, 1_000_000):
    m1 = torch.rand(n, D, device=dev) - 0.5; m2 = torch.rand(n, D, device=dev) - 0.5
    v1 = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1); v2 = v1.clone()
    for cols in (K + 1, (K + 1) // 32):
        idx = torch.randint(0, n, (B, cols), device=dev)
        for _ in range(3): crd.crd_fused_loss_grad(m1, m2, v1, v2, idx, T, Z, n, cols - 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): crd.crd_fused_loss_grad(m1, m2, v1, v2, idx, T, Z, n, cols - 1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gb = 2 * B * cols * D * 4 / 1e9
        print(f"n={n:8d} ({2*n*D*4/1e6:7.1f} MB both banks) cols={cols:6d}: {ms:.3f} ms  {gb/ms*1e3:8.1f} GB/s", flush=True)
