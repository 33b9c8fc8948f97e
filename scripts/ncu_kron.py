#!/usr/bin/env python
"""ncu target: one warm pass and one profiled pass of the three tcgen05 Kronecker kernels at one shape.

    ncu --set full --clock-control none --import-source on -k regex:kron_(fwd|wgrad|dgrad)_tc_kernel -s 3 -c 3 \
        -o gpurun_out/prof python scripts/ncu_kron.py 16384,128,128,256 [dropout]
Launch order per pass: kron_fwd_tc_kernel, kron_wgrad_tc_kernel, kron_dgrad_tc_kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_learning_b200.fusion import KronLinearState, kron_linear  # noqa: E402

v = [int(x) for x in sys.argv[1].split(",")]
B, dims, N = v[0], tuple(v[1:-1]), v[-1]
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
kk = 1
for d in dims:
    kk *= d + 1
fs = [torch.rand(B, d, device=dev, generator=gen).requires_grad_(True) for d in dims]
W = (torch.randn(N, kk, device=dev, generator=gen) / kk ** 0.5).requires_grad_(True)
st = KronLinearState(dims)
for _ in range(2):
    y = kron_linear(st, fs, W, None, drop_p=p, training=p > 0, seed=7)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
