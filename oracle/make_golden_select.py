"""Generate tests/golden/crdsel_*.npz by running the UNMODIFIED reference selection variant of CRD on CPU:
`MICCAI-2022/CL_utils/CRD_loss.py:127-175` (5-arg CRDLoss) over `CL_utils/memory_new.py:225-397`
(ContrastMemory_v3) and `CRD_loss.py:212-252` (ContrastLoss_v2).

Run in the build container only (needs /root/reference):   python oracle/make_golden_select.py
Shims (the reference hard-codes CUDA): `AliasMethod.cuda = no-op` (memory_new.py:235) and, while the reference runs,
`torch.Tensor.cuda = identity` (memory_new.py:311-357 call `.cuda()` on freshly built index tensors).
The numpy picks of the positive selection (:311-321) come from the GLOBAL numpy RNG: each step is preceded by
`np.random.seed(seed)` and the seed is stored, so the oracle / the CUDA module replay the same picks.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.make_golden import REF_ROOT, _save  # noqa: E402


def _import_reference():
    for name in ("CL_utils", "CL_utils.CRD_loss", "CL_utils.memory_new"):
        sys.modules.pop(name, None)
    root = os.path.join(REF_ROOT, "MICCAI-2022")
    sys.path.insert(0, root)
    try:
        loss_mod = importlib.import_module("CL_utils.CRD_loss")
        mem_mod = importlib.import_module("CL_utils.memory_new")
    finally:
        sys.path.remove(root)
    mem_mod.AliasMethod.cuda = lambda self: None
    return loss_mod, mem_mod


@contextlib.contextmanager
def _cuda_is_identity():
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def gen(loss_mod, name, *, B, s_dim, t_dim, D, P, K, P2, K2, n, mode, select_neg="True", sample_KD="False",
        epochs=(0.0, 0.4), seed=2019):
    torch.manual_seed(seed)
    opt = types.SimpleNamespace(s_dim=s_dim, t_dim=t_dim, feat_dim=D, nce_p=P, nce_p2=P2, nce_k=K, nce_k2=K2,
                                nce_t=0.07, nce_m=0.5, select_pos_pairs=True, select_neg_pairs=select_neg,
                                sample_KD=sample_KD, select_pos_mode=mode)
    with contextlib.redirect_stdout(io.StringIO()):
        mod = loss_mod.CRDLoss(opt, n)
    arrays = {f"init.{k}": v.clone() for k, v in mod.state_dict().items()}
    captured = {}
    mod.contrast.register_forward_hook(
        lambda m, i, o: captured.update(out_v1=o[0].detach().clone(), out_v2=o[1].detach().clone()))
    for s, epoch in enumerate(epochs):
        f_s = torch.randn(B, s_dim, requires_grad=True)
        f_t = torch.randn(B, t_dim, requires_grad=True)
        idx = torch.randperm(n)[:B]
        cidx = torch.randint(0, n, (B, K + P))
        cidx[:, 0] = idx
        mod.zero_grad()
        np_seed = 1000 + 17 * s
        np.random.seed(np_seed)
        with contextlib.redirect_stdout(io.StringIO()), _cuda_is_identity():
            loss = mod(epoch, f_s, f_t, idx, cidx)
        G = torch.ones_like(loss) if loss.dim() == 0 else torch.linspace(0.5, 1.5, loss.numel())
        (loss * G).sum().backward()
        p = f"step{s}."
        arrays.update({p + "epoch": np.array(epoch, dtype=np.float64), p + "np_seed": np.array(np_seed, dtype=np.int64),
                       p + "f_s": f_s, p + "f_t": f_t, p + "idx": idx, p + "contrast_idx": cidx,
                       p + "loss": loss.detach().reshape(-1), p + "G": G.reshape(-1),
                       p + "grad_f_s": f_s.grad.clone(), p + "grad_f_t": f_t.grad.clone(),
                       p + "out_v1": captured["out_v1"], p + "out_v2": captured["out_v2"],
                       p + "params": mod.contrast.params.clone(),
                       p + "memory_v1": mod.contrast.memory_v1.clone(),
                       p + "memory_v2": mod.contrast.memory_v2.clone()})
        for k, v in mod.named_parameters():
            arrays[p + "grad." + k] = v.grad.clone()
    _save(name, dict(B=B, s_dim=s_dim, t_dim=t_dim, D=D, P=P, K=K, P2=P2, K2=K2, n=n, mode=mode,
                     select_neg_pairs=select_neg, sample_KD=sample_KD, steps=len(epochs), T=0.07, momentum=0.5), arrays)


def main():
    torch.set_num_threads(1)
    loss_mod, _ = _import_reference()
    gen(loss_mod, "crdsel_random", B=6, s_dim=10, t_dim=12, D=16, P=8, K=24, P2=3, K2=10, n=120, mode="random")
    gen(loss_mod, "crdsel_hard_d128", B=8, s_dim=64, t_dim=48, D=128, P=40, K=200, P2=5, K2=64, n=300, mode="hard")
    gen(loss_mod, "crdsel_mid_d64", B=5, s_dim=20, t_dim=20, D=64, P=100, K=60, P2=10, K2=32, n=400, mode="mid")
    gen(loss_mod, "crdsel_curriculum", B=4, s_dim=9, t_dim=7, D=32, P=150, K=50, P2=6, K2=20, n=300, mode="curriculum",
        epochs=(0.2, 0.5, 0.9))
    gen(loss_mod, "crdsel_allneg", B=6, s_dim=10, t_dim=12, D=16, P=8, K=24, P2=3, K2=10, n=120, mode="hard",
        select_neg="False")
    gen(loss_mod, "crdsel_sampleKD", B=6, s_dim=10, t_dim=12, D=16, P=8, K=24, P2=3, K2=10, n=120, mode="random",
        sample_KD="True")


if __name__ == "__main__":
    main()
