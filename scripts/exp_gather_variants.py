#!/usr/bin/env python
"""Time K4 (fused loss+grad) at BASELINE config 2 for the MML_CRD_VARIANT set in the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_learning_b200 import crd
dev = torch.device("cuda:0")
B, D, K, n = 1024, 128, 16384, 1_000_000
g = torch.Generator(device=dev).manual_seed(0)
m1 = torch.rand(n, D, device=dev, generator=g) - 0.5; m2 = torch.rand(n, D, device=dev, generator=g) - 0.5
v1 = torch.nn.functional.normalize(torch.randn(B, D, device=dev, generator=g), dim=1); v2 = v1.clone()
idxs = [torch.randint(0, n, (B, K + 1), device=dev, generator=g) for _ in range(3)]
Z = torch.tensor([2.2e6, 2.2e6], device=dev)
for i in range(3): crd.crd_fused_loss_grad(m1, m2, v1, v2, idxs[i % 3], 0.07, Z, n, K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30): crd.crd_fused_loss_grad(m1, m2, v1, v2, idxs[i % 3], 0.07, Z, n, K)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
print(f"variant {os.environ.get('MML_CRD_VARIANT','0')}: {ms:.4f} ms  {2*B*(K+1)*D*4/ms/1e6:.1f} GB/s", flush=True)
