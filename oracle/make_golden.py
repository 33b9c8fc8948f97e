"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on CPU.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py
The fixtures travel to the GPU box; /root/reference does not.  The three shims
are the ones SURVEY.md §8c lists (the reference hard-codes CUDA and imports
uninstalled plotting/survival packages through `utils`):
  1. a stub `utils` module exposing only `init_max_weights` (utils.py:239-244);
  2. `torch.cuda.FloatTensor = torch.FloatTensor` (fusion.py:56-57);
  3. `AliasMethod.cuda = no-op` (CRD_criterion.py:17,125-127).
No reference source is copied: the modules are imported from where they lie.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _import_reference(tree: str):
    """Import (fusion, CRD_criterion, KD_loss) from one reference sub-tree."""
    for name in ("utils", "fusion", "KD_loss", "CL_utils", "CL_utils.CRD_criterion"):
        sys.modules.pop(name, None)
    stub = types.ModuleType("utils")

    def init_max_weights(module):           # semantics of utils.py:239-244
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
        import importlib
        fusion = importlib.import_module("fusion")
        crd = importlib.import_module("CL_utils.CRD_criterion")
        kd = importlib.import_module("KD_loss")
    finally:
        sys.path.remove(root)
    crd.AliasMethod.cuda = lambda self: None
    return fusion, crd, kd


def _np(t):
    return t.detach().cpu().numpy().copy()


def _save(name, cfg, arrays):
    arrays = {k: (v if isinstance(v, np.ndarray) else _np(v)) for k, v in arrays.items()}
    arrays["__config__"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + ".npz")
    np.savez(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


# ------------------------------------------------------------------ CRD --- #
def gen_crd(crd, name, *, B, s_dim, t_dim, D, K, n, steps=2, seed=2019):
    torch.manual_seed(seed)
    opt = types.SimpleNamespace(s_dim=s_dim, t_dim=t_dim, feat_dim=D, n_data=n,
                                nce_k=K, nce_t=0.07, nce_m=0.5)
    with contextlib.redirect_stdout(io.StringIO()):
        mod = crd.CRDLoss(opt)
    arrays = {f"init.{k}": v.clone() for k, v in mod.state_dict().items()}
    arrays["alias.prob"] = mod.contrast.multinomial.prob
    arrays["alias.alias"] = mod.contrast.multinomial.alias
    captured = {}
    mod.contrast.register_forward_hook(
        lambda m, i, o: captured.update(out_v1=o[0].detach().clone(), out_v2=o[1].detach().clone()))
    for s in range(steps):
        f_s = torch.randn(B, s_dim, requires_grad=True)
        f_t = torch.randn(B, t_dim, requires_grad=True)
        idx = torch.randperm(n)[:B]
        cidx = torch.randint(0, n, (B, K + 1))
        cidx[:, 0] = idx
        mod.zero_grad()
        with contextlib.redirect_stdout(io.StringIO()):
            loss = mod(f_s, f_t, idx, cidx)
        loss.backward()
        p = f"step{s}."
        arrays.update({p + "f_s": f_s, p + "f_t": f_t, p + "idx": idx, p + "contrast_idx": cidx,
                       p + "loss": loss, p + "grad_f_s": f_s.grad.clone(), p + "grad_f_t": f_t.grad.clone(),
                       p + "out_v1": captured["out_v1"], p + "out_v2": captured["out_v2"],
                       p + "params": mod.contrast.params.clone(),
                       p + "memory_v1": mod.contrast.memory_v1.clone(),
                       p + "memory_v2": mod.contrast.memory_v2.clone()})
        for k, v in mod.named_parameters():
            arrays[p + "grad." + k] = v.grad.clone()
    _save(name, dict(B=B, s_dim=s_dim, t_dim=t_dim, D=D, K=K, n=n, steps=steps,
                     T=0.07, momentum=0.5), arrays)


def gen_alias(crd):
    rng = np.random.default_rng(7)
    cases = {
        "nonuniform37": rng.random(37).astype(np.float32) * 3,
        "nonuniform1000": (rng.random(1000).astype(np.float32) ** 4) * 10,
        "uniform4096": np.ones(4096, dtype=np.float32),
        "uniform1000": np.ones(1000, dtype=np.float32),
        "sub_unit_sum": (rng.random(64).astype(np.float32) / 64),       # sum < 1: no normalisation (:90)
        "two_spikes": np.array([1e-3] * 30 + [5.0, 7.0], dtype=np.float32),
    }
    arrays, names = {}, []
    for cname, raw in cases.items():
        probs = torch.from_numpy(raw.copy())
        am = crd.AliasMethod(probs)            # mutates `probs` in place when sum > 1 (:90-91)
        arrays[f"{cname}.raw"] = raw
        arrays[f"{cname}.normalised"] = probs
        arrays[f"{cname}.prob"] = am.prob
        arrays[f"{cname}.alias"] = am.alias
        # draw: reference output, then the same raw draws replayed call-for-call (:133,137)
        N = 2000
        torch.manual_seed(11)
        out = am.draw(N)
        torch.manual_seed(11)
        kk = torch.zeros(N, dtype=torch.long).random_(0, len(raw))
        b = torch.bernoulli(am.prob.index_select(0, kk))
        arrays[f"{cname}.draw_kk"] = kk
        arrays[f"{cname}.draw_b"] = b
        arrays[f"{cname}.draw_out"] = out
        names.append(cname)
    _save("alias", dict(cases=names, draw_n=2000), arrays)


# --------------------------------------------------------------- fusion --- #
def _run_fusion(mod, inputs, training, arrays, tag):
    mod.train(training)
    ins = [x.clone().requires_grad_(True) for x in inputs]
    mod.zero_grad()
    out = mod(*ins)
    torch.manual_seed(99)
    G = torch.randn_like(out)
    (out * G).sum().backward()
    arrays[f"{tag}.out"] = out
    arrays[f"{tag}.G"] = G
    for i, x in enumerate(ins):
        arrays[f"{tag}.grad_vec{i + 1}"] = x.grad
    for k, v in mod.named_parameters():
        arrays[f"{tag}.grad.{k}"] = v.grad if v.grad is not None else torch.zeros_like(v)
    for k, v in mod.state_dict().items():
        if "running" in k or "num_batches" in k:
            arrays[f"{tag}.after.{k}"] = v.clone()


def gen_bilinear(fusion, name, *, B, seed=2019, train_mode=True, **kw):
    torch.manual_seed(seed)
    kw = dict(kw)
    kw.setdefault("dropout_rate", 0.0)       # p=0: train-mode BN stats without torch's mask stream
    mod = fusion.BilinearFusion(**kw)
    # non-trivial BN affine + running stats so eval mode exercises them
    for bn in (mod.encoder1[1], mod.encoder2[1]):
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.normal_(0, 0.2)
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 2.0)
    arrays = {f"init.{k}": v.clone() for k, v in mod.state_dict().items()}
    inputs = [torch.randn(B, kw.get("dim1", 32)), torch.randn(B, kw.get("dim2", 32))]
    arrays["vec1"], arrays["vec2"] = inputs
    _run_fusion(mod, inputs, False, arrays, "eval")
    if train_mode:
        _run_fusion(mod, inputs, True, arrays, "train")
    _save(name, dict(B=B, kind="bilinear", **kw), arrays)


def gen_trilinear(fusion, name, variant, *, B, seed=2019, **kw):
    torch.manual_seed(seed)
    cls = fusion.TrilinearFusion_A if variant == "A" else fusion.TrilinearFusion_B
    mod = cls(**kw)
    arrays = {f"init.{k}": v.clone() for k, v in mod.state_dict().items()}
    inputs = [torch.randn(B, kw["dim1"]), torch.randn(B, kw["dim2"]), torch.randn(B, kw["dim3"])]
    arrays["vec1"], arrays["vec2"], arrays["vec3"] = inputs
    _run_fusion(mod, inputs, False, arrays, "eval")   # post_fusion_dropout is a hard-coded p=0.25 (:93)
    _save(name, dict(B=B, kind="trilinear", variant=variant, **kw), arrays)


def gen_kd(kd):
    torch.manual_seed(5)
    arrays = {}
    for i, (B, C, T) in enumerate([(16, 3, 4.0), (7, 10, 1.0), (64, 3, 2.5)]):
        y_s = torch.randn(B, C, requires_grad=True)
        y_t = torch.randn(B, C)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            loss = kd.DistillKL(T)(y_s, y_t)
        loss.backward()
        arrays.update({f"c{i}.y_s": y_s, f"c{i}.y_t": y_t, f"c{i}.T": np.array(T, dtype=np.float64),
                       f"c{i}.loss": loss, f"c{i}.grad_y_s": y_s.grad})
    _save("distill_kl", dict(cases=3), arrays)


def main():
    ap = argparse.ArgumentParser()
    ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)                  # deterministic reduction order
    fusion, crd, kd = _import_reference("MICCAI-2022")
    gen_crd(crd, "crd_small", B=8, s_dim=12, t_dim=10, D=16, K=32, n=100)
    gen_crd(crd, "crd_d128", B=16, s_dim=64, t_dim=64, D=128, K=128, n=300)
    gen_crd(crd, "crd_d64_ragged", B=5, s_dim=7, t_dim=9, D=64, K=33, n=61, steps=3)
    gen_alias(crd)
    gen_bilinear(fusion, "bilinear_c1", B=16, train_mode=False,
                 skip=0, dim1=32, dim2=32, mmhid=64)
    gen_bilinear(fusion, "bilinear_skip", B=12, skip=1, dim1=16, dim2=16, mmhid=32)
    gen_bilinear(fusion, "bilinear_odd", B=9, skip=1, use_bilinear=0, gate1=1, gate2=0,
                 dim1=8, dim2=12, mmhid=16)
    gen_bilinear(fusion, "bilinear_scaled", B=10, skip=0, dim1=16, dim2=24, scale_dim1=2,
                 scale_dim2=3, mmhid=24)
    gen_trilinear(fusion, "trilinear_A", "A", B=8, skip=1, dim1=8, dim2=6, dim3=10, mmhid=24)
    gen_trilinear(fusion, "trilinear_B", "B", B=8, skip=0, dim1=12, dim2=12, dim3=12, mmhid=16,
                  gate2=0)
    gen_kd(kd)
    # the MIA-2022 tree's single-Linear Embed variant (MIA 2022/CL_utils/CRD_criterion.py:223)
    _, crd2, _ = _import_reference("MIA 2022")
    gen_crd(crd2, "crd_embed1", B=4, s_dim=6, t_dim=5, D=16, K=8, n=50)


if __name__ == "__main__":
    main()
