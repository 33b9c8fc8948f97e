#!/usr/bin/env python
"""Observed max relative errors (conftest.rel_err) of every fusion golden fixture, per kernel path and mode: the numbers
behind the tolerances in tests/test_fusion_gpu.py.   python scripts/report_fusion_errors.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import multimodal_learning_b200 as pkg  # noqa: E402
from conftest import Golden, rel_err  # noqa: E402
from test_fusion_gpu import GOLDENS, _make  # noqa: E402

DEV = "cuda:0"
for name in GOLDENS:
    g = Golden(name)
    nvec = 3 if g.cfg["kind"] == "trilinear" else 2
    for path in ("simt", "auto"):
        for tag in ["eval"] + (["train"] if any(k.startswith("train.") for k in g.keys()) else []):
            mod = _make(pkg, g)
            mod.set_kron_path(path)
            mod.train(tag == "train")
            ins = [g.t(f"vec{i + 1}", DEV).requires_grad_(True) for i in range(nvec)]
            out = mod(*ins)
            (out * g.t(f"{tag}.G", DEV)).sum().backward()
            errs = {"out": rel_err(out, g.t(f"{tag}.out"))}
            for i, x in enumerate(ins):
                errs[f"vec{i + 1}"] = rel_err(x.grad, g.t(f"{tag}.grad_vec{i + 1}"))
            for k, v in mod.named_parameters():
                want = g.t(f"{tag}.grad.{k}")
                if want.abs().max() >= 1e-4:
                    errs[k] = rel_err(v.grad if v.grad is not None else torch.zeros_like(v), want)
            worst = max(errs, key=errs.get)
            print(f"{name:16s} {path:5s} {tag:5s} out {errs['out']:.2e}  worst {worst} {errs[worst]:.2e}", flush=True)
