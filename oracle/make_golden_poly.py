"""Generate tests/golden/polynomial_*.npz by running the UNMODIFIED `PolynomialFusion`
(`MIA 2023/stage2_unimodal_student/fusion.py:6-73`) on CPU with the shims of oracle/make_golden.py.

Run in the build container only (needs /root/reference):   python oracle/make_golden_poly.py
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import importlib
import math
import types

import torch.nn as nn

from oracle.make_golden import REF_ROOT, _run_fusion, _save  # noqa: E402


def gen(fusion, name, *, B, seed=2019, **kw):
    torch.manual_seed(seed)
    kw = dict(kw)
    kw.setdefault("dropout_rate", 0.0)       # p=0: train-mode BN statistics without torch's mask stream
    mod = fusion.PolynomialFusion(**kw)
    for bn in (mod.encoder1[1], mod.encoder2[1], mod.encoder3[1]):
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.normal_(0, 0.2)
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 2.0)
    arrays = {f"init.{k}": v.clone() for k, v in mod.state_dict().items()}
    inputs = [torch.randn(B, kw.get("dim1", 32)), torch.randn(B, kw.get("dim2", 32))]
    arrays["vec1"], arrays["vec2"] = inputs
    _run_fusion(mod, inputs, False, arrays, "eval")
    _run_fusion(mod, inputs, True, arrays, "train")
    _save(name, dict(B=B, kind="polynomial", **kw), arrays)


def _import_fusion(tree):
    """`fusion` of one reference sub-tree with the `utils` stub (utils.py:239-244) and the CPU FloatTensor shim."""
    for name in ("utils", "fusion"):
        sys.modules.pop(name, None)
    stub = types.ModuleType("utils")

    def init_max_weights(module):
        for m in module.modules():
            if type(m) == nn.Linear:
                stdv = 1.0 / math.sqrt(m.weight.size(1))
                m.weight.data.normal_(0, stdv)
                m.bias.data.zero_()
    stub.init_max_weights = init_max_weights
    sys.modules["utils"] = stub
    torch.cuda.FloatTensor = torch.FloatTensor
    root = os.path.join(REF_ROOT, tree)
    sys.path.insert(0, root)
    try:
        return importlib.import_module("fusion")
    finally:
        sys.path.remove(root)


def main():
    torch.set_num_threads(1)
    fusion = _import_fusion("MIA 2023/stage2_unimodal_student")
    gen(fusion, "polynomial_16", B=12, skip=1, dim1=16, dim2=16, mmhid=16)        # (mmhid+1)^2 must equal (dim1+1)(dim2+1)
    gen(fusion, "polynomial_gate", B=9, skip=0, use_bilinear=0, gate1=0, gate2=1, dim1=8, dim2=8, mmhid=8)


if __name__ == "__main__":
    main()
