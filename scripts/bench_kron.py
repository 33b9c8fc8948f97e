#!/usr/bin/env python
"""Kronecker-fusion kernel sweep (BASELINE configs 3 and 4): achieved TFLOP/s of the tcgen05 forward
(and the time of the fp32 backward) per shape, CUDA-event timed.  One JSON line per shape.

    python scripts/bench_kron.py [--quick] [--only B,d,N] [--iters 30]
FLOPs are algorithmic: 2*B*Kk*N forward, Kk = prod(d_i + 1)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def time_ms(fn, iters, warmup=5):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None, help="B,d,N  or  B,d,d,d,N")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--bwd", action="store_true", help="also time the backward kernels")
    ap.add_argument("--dropout", type=float, default=0.0)
    args = ap.parse_args()
    import multimodal_learning_b200 as pkg
    from multimodal_learning_b200.fusion import KronLinearState, kron_linear
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
    tf32_peak = peaks["bf16_tflops"] / 2
    if args.only:
        v = [int(x) for x in args.only.split(",")]
        shapes = [(v[0], tuple(v[1:-1]), v[-1])]
    elif args.quick:
        shapes = [(4096, (32, 32), 64), (65536, (32, 32), 256), (16384, (128, 128), 256), (8192, (32, 32, 32), 96)]
    else:
        shapes = [(B, (d, d), N) for d in (32, 64, 128) for N in (64, 128, 256) for B in (4096, 16384, 65536)]
        shapes.append((8192, (32, 32, 32), 96))
    for B, dims, N in shapes:
        kk = 1
        for d in dims:
            kk *= d + 1
        gen = torch.Generator(device=dev).manual_seed(0)
        fs = [torch.rand(B, d, device=dev, generator=gen) for d in dims]
        W = torch.randn(N, kk, device=dev, generator=gen) / kk ** 0.5
        bias = torch.zeros(N, device=dev)
        st = KronLinearState(dims)
        training = args.dropout > 0
        fwd = lambda: kron_linear(st, fs, W, bias, drop_p=args.dropout, training=training, seed=7)
        ms = time_ms(fwd, args.iters)
        flops = 2.0 * B * kk * N
        line = {"B": B, "dims": dims, "N": N, "Kk": kk, "fwd_ms": round(ms, 4), "fwd_tflops": round(flops / ms / 1e9, 1),
                "frac_of_tf32_peak": round(flops / ms / 1e9 / tf32_peak, 3), "tf32_peak": tf32_peak,
                "dropout": args.dropout}
        if args.bwd:
            # `_KronLinearFn.backward` runs the kernels of whichever inputs required grad at FORWARD time, so three
            # separate graphs time (wgrad + dgrad), wgrad alone (K3 + transpose + unpack) and dgrad alone (K2 + reduce).
            n_b = max(3, args.iters // 3)
            fsg = [f.clone().requires_grad_(True) for f in fs]
            Wg = W.detach().clone().requires_grad_(True)
            Wc = W.detach()
            y_all = kron_linear(st, fsg, Wg, bias, drop_p=args.dropout, training=training, seed=7)
            y_w = kron_linear(st, fs, Wg, bias, drop_p=args.dropout, training=training, seed=7)
            y_f = kron_linear(st, fsg, Wc, bias, drop_p=args.dropout, training=training, seed=7)
            G = torch.randn_like(y_all)
            line["bwd_ms"] = round(time_ms(lambda: torch.autograd.grad(y_all, [Wg] + fsg, G, retain_graph=True), n_b, warmup=2), 4)
            line["wgrad_ms"] = round(time_ms(lambda: torch.autograd.grad(y_w, [Wg], G, retain_graph=True), n_b, warmup=2), 4)
            line["dgrad_ms"] = round(time_ms(lambda: torch.autograd.grad(y_f, fsg, G, retain_graph=True), n_b, warmup=2), 4)
            line["bwd_tflops"] = round(2 * flops / line["bwd_ms"] / 1e9, 1)
            line["wgrad_tflops"] = round(flops / line["wgrad_ms"] / 1e9, 1)
            line["dgrad_tflops"] = round(flops / line["dgrad_ms"] / 1e9, 1)
            line["bwd_frac_of_tf32_peak"] = round(line["bwd_tflops"] / tf32_peak, 3)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
