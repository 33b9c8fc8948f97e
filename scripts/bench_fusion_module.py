#!/usr/bin/env python
"""Whole-module timing of BilinearFusion / TrilinearFusion_A forward+backward (gates + Kronecker encoder)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_learning_b200 as pkg

def time_ms(fn, iters=10, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

dev = torch.device("cuda:0")
for B, d, N in [(64, 32, 64), (4096, 32, 64), (16384, 64, 128), (16384, 128, 256)]:
    mod = pkg.BilinearFusion(skip=0, dim1=d, dim2=d, mmhid=N, dropout_rate=0.25).to(dev).train()
    v1 = torch.randn(B, d, device=dev, requires_grad=True); v2 = torch.randn(B, d, device=dev, requires_grad=True)
    def step():
        mod.zero_grad(set_to_none=True)
        out = mod(v1, v2); out.sum().backward()
    ms = time_ms(step)
    kk = (d + 1) ** 2
    print(json.dumps({"module": "BilinearFusion", "B": B, "d": d, "N": N, "train_fwd_bwd_ms": round(ms, 3),
                      "encoder1_tflops": round(6 * B * kk * N / ms / 1e9, 1),
                      "gates_tflops": round(2 * 6 * B * d * d * d / ms / 1e9, 1)}), flush=True)
mod = pkg.TrilinearFusion_A(skip=1, dim1=32, dim2=32, dim3=32, mmhid=96).to(dev).train()
B = 8192
vs = [torch.randn(B, 32, device=dev, requires_grad=True) for _ in range(3)]
def step3():
    mod.zero_grad(set_to_none=True)
    mod(*vs).sum().backward()
ms = time_ms(step3)
print(json.dumps({"module": "TrilinearFusion_A", "B": B, "d": 32, "N": 96, "train_fwd_bwd_ms": round(ms, 3),
                  "encoder1_tflops": round(6 * B * 33 ** 3 * 96 / ms / 1e9, 1)}), flush=True)
