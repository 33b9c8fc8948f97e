"""Generate tests/golden/crdknn_*.npz by running the UNMODIFIED reference KNN / class-centre CRD on CPU:
`MIA 2023/stage2_unimodal_student/CL_utils/CRD_criterion_v10.py` (`CRDLoss(opt, n_data, train_class_idx)`).

Run in the build container only (needs /root/reference and scikit-learn):   python oracle/make_golden_knn.py
Shim: `torch.Tensor.cuda = identity` while the reference runs (it hard-codes `.cuda()` at :56-76)."""
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
    root = os.path.join(REF_ROOT, "MIA 2023", "stage2_unimodal_student")
    sys.path.insert(0, root)
    try:
        return importlib.import_module("CL_utils.CRD_criterion_v10")
    finally:
        sys.path.remove(root)


class _RecordingKMeans:
    """Stands in for `KMeans` inside the reference module (CRD_criterion_v10.py:89-92): draws the k-means++ initial centres
    itself (seeded), runs the real sklearn estimator from them, and keeps (init, centres) of every fit in call order."""
    log = []
    rng = np.random.RandomState(0)

    def __init__(self, n_clusters):
        self.k = n_clusters

    def fit(self, X):
        from sklearn.cluster import KMeans, kmeans_plusplus
        init, _ = kmeans_plusplus(X, self.k, random_state=_RecordingKMeans.rng)
        est = KMeans(n_clusters=self.k, init=init.copy(), n_init=1).fit(X)
        self.cluster_centers_ = est.cluster_centers_
        _RecordingKMeans.log.append((init.astype(np.float32), est.cluster_centers_.astype(np.float32), est.n_iter_))
        return self


def gen(mod_v10, name, *, pos_extra, B, s_dim, t_dim, D, P, K, n, steps=2, seed=2023):
    torch.manual_seed(seed)
    record = pos_extra == "centers" and P > 2
    if record:
        mod_v10.KMeans = _RecordingKMeans
        _RecordingKMeans.rng = np.random.RandomState(seed)
    rng = np.random.default_rng(seed)
    cls = rng.integers(0, 3, size=n)
    class_idx = [np.nonzero(cls == c)[0] for c in range(3)]
    opt = types.SimpleNamespace(s_dim=s_dim, t_dim=t_dim, feat_dim=D, nce_k=K, nce_t=0.07, nce_m=0.5, nce_p=P, pos_extra=pos_extra)
    with contextlib.redirect_stdout(io.StringIO()):
        mod = mod_v10.CRDLoss(opt, n, class_idx)
    arrays = {f"init.{k}": v.clone() for k, v in mod.state_dict().items()}
    arrays["row_class"] = torch.as_tensor(cls)
    captured = {}
    mod.contrast.register_forward_hook(lambda m, i, o: captured.update(out=[t.detach().clone() for t in o]))
    for s in range(steps):
        f_s = torch.randn(B, s_dim, requires_grad=True)
        f_t = torch.randn(B, t_dim, requires_grad=True)
        idx = torch.randperm(n)[:B]
        label = torch.as_tensor(cls)[idx].long()
        cidx = torch.randint(0, n, (B, K + 1))
        cidx[:, 0] = idx
        w = torch.rand(B) + 0.5
        mod.zero_grad()
        _RecordingKMeans.log.clear()
        with contextlib.redirect_stdout(io.StringIO()), _cuda_is_identity():
            loss, sample_loss = mod(w, f_s, f_t, label, idx, cidx)
        loss.backward()
        p = f"step{s}."
        arrays.update({p + "f_s": f_s, p + "f_t": f_t, p + "idx": idx, p + "label": label, p + "contrast_idx": cidx,
                       p + "sample_weights": w, p + "loss": loss.detach().reshape(-1), p + "sample_loss": sample_loss.detach(),
                       p + "grad_f_s": f_s.grad.clone(), p + "grad_f_t": f_t.grad.clone(),
                       p + "out_v1": captured["out"][0], p + "out_v2": captured["out"][1],
                       p + "params": mod.contrast.params.clone(),
                       p + "memory_v1": mod.contrast.memory_v1.clone(), p + "memory_v2": mod.contrast.memory_v2.clone()})
        if pos_extra == "neighbors":
            arrays[p + "sim_v1"], arrays[p + "sim_v2"] = captured["out"][2], captured["out"][3]
        if record:                                           # fits in call order: bank 1 classes 0..2, then bank 2 classes 0..2
            log = _RecordingKMeans.log
            assert len(log) == 6
            arrays[p + "kmeans_init"] = torch.from_numpy(np.stack([e[0] for e in log]).reshape(2, 3, P - 1, D))
            arrays[p + "kmeans_centres"] = torch.from_numpy(np.stack([e[1] for e in log]).reshape(2, 3, P - 1, D))
            arrays[p + "kmeans_iters"] = torch.tensor([e[2] for e in log])
        for k, v in mod.named_parameters():
            arrays[p + "grad." + k] = v.grad.clone()
    _save(name, dict(pos_extra=pos_extra, B=B, s_dim=s_dim, t_dim=t_dim, D=D, P=P, K=K, n=n, steps=steps, T=0.07, momentum=0.5),
          arrays)


def main():
    torch.set_num_threads(1)
    m = _import_reference()
    gen(m, "crdknn_p3_d16", pos_extra="neighbors", B=6, s_dim=10, t_dim=12, D=16, P=3, K=20, n=150)
    gen(m, "crdknn_p5_d128", pos_extra="neighbors", B=8, s_dim=24, t_dim=20, D=128, P=5, K=40, n=300, steps=3)
    gen(m, "crdknn_p1_d32", pos_extra="neighbors", B=5, s_dim=10, t_dim=12, D=32, P=1, K=20, n=140)
    gen(m, "crdknn_centers_d32", pos_extra="centers", B=6, s_dim=10, t_dim=12, D=32, P=2, K=20, n=150)
    gen(m, "crdknn_kmeans_p4_d32", pos_extra="centers", B=6, s_dim=10, t_dim=12, D=32, P=4, K=20, n=600)
    gen(m, "crdknn_kmeans_p3_d128", pos_extra="centers", B=8, s_dim=16, t_dim=12, D=128, P=3, K=30, n=480)
    gen(m, "crdknn_kmeans_p6_d64", pos_extra="centers", B=6, s_dim=12, t_dim=10, D=64, P=6, K=20, n=450)


if __name__ == "__main__":
    main()
