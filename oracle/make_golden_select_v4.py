"""Generate tests/golden/crdv4_*.npz and crdmono_*.npz by running the UNMODIFIED reference classes of the MIA 2022 tree on
CPU: `MIA 2022/CL_utils/CRD_loss_v2.py:13-55` CRDLoss over `CL_utils/memory_new.py:398-561` ContrastMemory_v4, and
`CRD_loss_v2.py:58-104` CRDLoss_v2 over `memory_new.py:565-698` ContrastMemory_mono.

Run in the build container only (needs /root/reference):   python oracle/make_golden_select_v4.py
Shims as in make_golden_select.py: `AliasMethod.cuda = no-op` (memory_new.py:408, :575) and, while the reference runs,
`torch.Tensor.cuda = identity`; every step is preceded by `np.random.seed(seed)` (stored) because the positive picks
come from the global numpy RNG.
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
from oracle.make_golden_select import _cuda_is_identity  # noqa: E402


def _import_reference():
    for name in [m for m in sys.modules if m == "CL_utils" or m.startswith("CL_utils.")]:
        sys.modules.pop(name)
    root = os.path.join(REF_ROOT, "MIA 2022")
    sys.path.insert(0, root)
    try:
        loss_mod = importlib.import_module("CL_utils.CRD_loss_v2")
        mem_mod = importlib.import_module("CL_utils.memory_new")
    finally:
        sys.path.remove(root)
    mem_mod.AliasMethod.cuda = lambda self: None
    return loss_mod


def gen(loss_mod, name, *, kind, B, s_dim, t_dim, D, P, K, P2, n, mode, neg_reweight="True", sample_KD="False",
        epochs=(0.0, 0.4), seed=2022):
    torch.manual_seed(seed)
    opt = types.SimpleNamespace(s_dim=s_dim, t_dim=t_dim, feat_dim=D, nce_p=P, nce_p2=P2, nce_k=K, nce_k2=K,
                                nce_t=0.07, nce_m=0.5, select_pos_pairs=True, select_neg_pairs="False",
                                neg_reweight=neg_reweight, sample_KD=sample_KD, select_pos_mode=mode)
    with contextlib.redirect_stdout(io.StringIO()):
        mod = (loss_mod.CRDLoss if kind == "v4" else loss_mod.CRDLoss_v2)(opt, n)
    arrays = {f"init.{k}": v.clone() for k, v in mod.state_dict().items()}
    captured = {}

    def hook(m, i, o):
        if kind == "v4":
            captured.update(out_v1=o[0].detach().clone(), out_v2=o[1].detach().clone())
        else:
            captured.update(out_v2=o[0].detach().clone())
    mod.contrast.register_forward_hook(hook)
    for s, epoch in enumerate(epochs):
        f_s = torch.randn(B, s_dim, requires_grad=True)
        f_t = torch.randn(B, t_dim, requires_grad=True)
        idx = torch.randperm(n)[:B]
        cidx = torch.randint(0, n, (B, K + P))
        cidx[:, 0] = idx
        mod.zero_grad()
        np_seed = 2000 + 13 * s
        np.random.seed(np_seed)
        with contextlib.redirect_stdout(io.StringIO()), _cuda_is_identity():
            loss = mod(epoch, f_s, f_t, idx, cidx)
        G = torch.ones_like(loss) if loss.dim() == 0 else torch.linspace(0.5, 1.5, loss.numel())
        (loss * G).sum().backward()
        p = f"step{s}."
        arrays.update({p + "epoch": np.array(epoch, dtype=np.float64), p + "np_seed": np.array(np_seed, dtype=np.int64),
                       p + "f_s": f_s, p + "f_t": f_t, p + "idx": idx, p + "contrast_idx": cidx,
                       p + "loss": loss.detach().reshape(-1), p + "G": G.reshape(-1),
                       p + "grad_f_s": f_s.grad.clone(),
                       p + "params": mod.contrast.params.clone(),
                       p + "memory_v1": mod.contrast.memory_v1.clone(),
                       p + "memory_v2": mod.contrast.memory_v2.clone()})
        if kind == "v4":
            arrays[p + "grad_f_t"] = f_t.grad.clone()
        else:
            assert f_t.grad is None                # CRD_loss_v2.py:92: the teacher feature is detached
        for k, v in captured.items():
            arrays[p + k] = v
        for k, v in mod.named_parameters():
            arrays[p + "grad." + k] = v.grad.clone()
    _save(name, dict(kind=kind, B=B, s_dim=s_dim, t_dim=t_dim, D=D, P=P, K=K, P2=P2, n=n, mode=mode,
                     neg_reweight=neg_reweight, sample_KD=sample_KD, steps=len(epochs), T=0.07, momentum=0.5), arrays)


def main():
    torch.set_num_threads(1)
    loss_mod = _import_reference()
    gen(loss_mod, "crdv4_hard", kind="v4", B=6, s_dim=10, t_dim=12, D=16, P=8, K=24, P2=3, n=120, mode="hard")
    gen(loss_mod, "crdv4_mid_d128", kind="v4", B=5, s_dim=40, t_dim=24, D=128, P=100, K=150, P2=10, n=260, mode="mid")
    gen(loss_mod, "crdv4_plain", kind="v4", B=6, s_dim=10, t_dim=12, D=32, P=8, K=24, P2=3, n=120, mode="random",
        neg_reweight="False")
    gen(loss_mod, "crdv4_curriculum_KD", kind="v4", B=4, s_dim=9, t_dim=7, D=32, P=150, K=50, P2=6, n=300,
        mode="curriculum", sample_KD="True", epochs=(0.2, 0.5, 0.9))
    # mono: the teacher feature already has feat_dim columns (t_dim == D)
    gen(loss_mod, "crdmono_hard", kind="mono", B=6, s_dim=10, t_dim=16, D=16, P=8, K=24, P2=3, n=120, mode="hard")
    gen(loss_mod, "crdmono_mid_d128", kind="mono", B=5, s_dim=40, t_dim=128, D=128, P=100, K=150, P2=10, n=260, mode="mid")
    gen(loss_mod, "crdmono_random_KD", kind="mono", B=6, s_dim=10, t_dim=32, D=32, P=8, K=24, P2=3, n=120, mode="random",
        sample_KD="True", epochs=(0.0, 0.3, 0.8))


if __name__ == "__main__":
    main()
