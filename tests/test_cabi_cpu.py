"""CPU checks of the C-ABI boundary: the library builds/loads, exports every symbol the
header declares, rejects bad arguments, and its HOST function (alias tables) is bit-exact
against the reference's golden vectors.  No GPU compute is issued here."""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np
import pytest
import torch

import multimodal_learning_b200 as pkg
from multimodal_learning_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_in_header():
    text = open(os.path.join(ROOT, "include", "mml_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mml_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_in_header()
    assert len(names) >= 10
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/mml_b200.h but not exported"
    # and the ctypes binding covers exactly the header
    assert names == _cabi.declared_symbols()


def test_abi_version_and_error_string():
    lib = _cabi.lib()
    assert lib.mml_abi_version() == _cabi.ABI_VERSION == 3
    rc = lib.mml_alias_build_host(None, 0, None, None)
    assert rc == -1
    assert b"alias_build" in lib.mml_last_error()
    assert lib.mml_crd_workspace_bytes(1024, 16385, 128) > 1024 * 2 * 128 * 4


def test_kernels_refuse_cpu_tensors():
    with pytest.raises(RuntimeError):
        _cabi.dptr(torch.zeros(4))
    am = pkg.AliasMethod(torch.ones(16))
    with pytest.raises(RuntimeError):
        am.draw(8)                       # no CPU fallback


def test_class_rows_host_logic_and_kmeans_refuses_cpu_tensors():
    """The class row lists of the k-means centres (crd_kmeans.ClassRows) and the limits the C ABI reports, without a GPU."""
    km = pkg.crd_kmeans
    cls = km.ClassRows([[4, 1, 9], np.array([0, 2]), torch.tensor([3])], "cpu")
    assert cls.n_classes == 3 and cls.sizes == [3, 2, 1]
    assert cls.offsets.tolist() == [0, 3, 5, 6] and cls.offsets.dtype == torch.int64
    assert cls.rows.tolist() == [4, 1, 9, 0, 2, 3]
    with pytest.raises(RuntimeError):
        km.ClassRows([[0, 1], []], "cpu")                                 # a class without rows
    shard = km.ClassRows([[0, 1], []], "cpu", allow_empty=True)         # ... unless it lives on another shard
    assert shard.sizes == [2, 0] and shard.offsets.tolist() == [0, 2, 2]
    with pytest.raises(RuntimeError):
        km.ClassRows([[], []], "cpu", allow_empty=True)
    lib = _cabi.lib()
    assert lib.mml_crd_kmeans_max_clusters() == 8
    assert lib.mml_crd_kmeans_workspace_bytes(3, 3, 128) > 0
    assert lib.mml_crd_kmeans_workspace_bytes(3, 9, 128) == 0 and lib.mml_crd_kmeans_workspace_bytes(33, 3, 128) == 0
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        km.lloyd(torch.zeros(6, 32), cls, torch.zeros(3, 2, 32))         # no CPU fallback


def test_alias_build_host_bit_exact(golden):
    g = golden("alias")
    for c in g.cfg["cases"]:
        raw = torch.from_numpy(g.np(f"{c}.raw").copy())
        am = pkg.AliasMethod(raw)        # normalises in place like the reference (:90-91)
        assert np.array_equal(raw.numpy().view(np.uint32), g.np(f"{c}.normalised").view(np.uint32)), c
        assert np.array_equal(am.prob.numpy().view(np.uint32), g.np(f"{c}.prob").view(np.uint32)), c
        assert np.array_equal(am.alias.numpy(), g.np(f"{c}.alias")), c
        assert am.alias.dtype == torch.int64 and am.prob.dtype == torch.float32


def test_alias_build_host_matches_oracle_random():
    from oracle import crd_oracle as co
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 17, 256, 5000):
        p = torch.from_numpy(rng.random(n).astype(np.float32) ** 3 + 1e-6)
        am = pkg.AliasMethod(p)
        prob, alias = co.alias_build(p.numpy())
        assert np.array_equal(am.prob.numpy().view(np.uint32), prob.view(np.uint32)), n
        assert np.array_equal(am.alias.numpy(), alias), n


def test_alias_uniform_large_is_identity():
    am = pkg.AliasMethod(torch.ones(1 << 20))
    assert bool((am.prob == 1).all()) and bool((am.alias == 0).all())


def test_module_state_dict_layout_matches_reference(golden):
    """Same seed -> same initial state as the reference (RNG call order preserved),
    and identical state_dict keys/shapes so checkpoints are interchangeable."""
    import types
    g = golden("crd_small")
    c = g.cfg
    torch.manual_seed(2019)
    opt = types.SimpleNamespace(s_dim=c["s_dim"], t_dim=c["t_dim"], feat_dim=c["D"], n_data=c["n"],
                                nce_k=c["K"], nce_t=0.07, nce_m=0.5)
    mod = pkg.CRDLoss(opt)
    want = g.state_dict("init.")
    got = mod.state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert got[k].shape == want[k].shape and got[k].dtype == want[k].dtype, k
        assert torch.equal(got[k], want[k]), k
    assert [n for n, _ in mod.named_parameters()] == [
        "embed_s.linear.0.weight", "embed_s.linear.0.bias", "embed_s.linear.2.weight", "embed_s.linear.2.bias",
        "embed_t.linear.0.weight", "embed_t.linear.0.bias", "embed_t.linear.2.weight", "embed_t.linear.2.bias"]


def test_distill_kl_matches_golden(golden):
    g = golden("distill_kl")
    for i in range(g.cfg["cases"]):
        y_s = g.t(f"c{i}.y_s").requires_grad_(True)
        loss = pkg.DistillKL(float(g.np(f"c{i}.T")))(y_s, g.t(f"c{i}.y_t"))
        loss.backward()
        assert abs(loss.item() - g.t(f"c{i}.loss").item()) < 1e-6 * max(1, abs(loss.item()))
        assert torch.allclose(y_s.grad, g.t(f"c{i}.grad_y_s"), rtol=1e-5, atol=1e-7)


def test_contrast_loss_standalone_matches_oracle():
    from oracle import crd_oracle as co
    torch.manual_seed(0)
    x = torch.rand(6, 9, 1) * 1e-2
    got = pkg.ContrastLoss(100)(x)
    assert got.shape == (1,)
    assert torch.allclose(got, co.nce_loss(x, 100), rtol=1e-6)


def test_selection_variant_state_dict_layout_matches_reference(golden):
    """5-arg CRDLoss(opt, n_data) (CL_utils/CRD_loss.py:133-151): same seed -> same initial state, same keys/shapes
    (params has 6 entries [K, T, Z_v1, Z_v2, momentum, P]; single-Linear Embed heads)."""
    import types
    g = golden("crdsel_random")
    c = g.cfg
    torch.manual_seed(2019)
    opt = types.SimpleNamespace(s_dim=c["s_dim"], t_dim=c["t_dim"], feat_dim=c["D"], nce_p=c["P"], nce_p2=c["P2"], nce_k=c["K"],
                                nce_k2=c["K2"], nce_t=0.07, nce_m=0.5, select_pos_pairs=True, select_neg_pairs="True",
                                sample_KD="False", select_pos_mode=c["mode"])
    mod = pkg.crd_select.CRDLoss(opt, c["n"])
    want, got = g.state_dict("init."), mod.state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert got[k].shape == want[k].shape and torch.equal(got[k], want[k]), k
    assert mod.contrast.params.numel() == 6
    # the criterion modules on host tensors (pure elementwise torch) agree with the oracle restatement
    from oracle import crd_select_oracle as so
    x = torch.rand(5, 3 + 7, 1) * 1e-2
    for kd in ("False", "True"):
        assert torch.allclose(pkg.ContrastLoss_v2(50, kd)(x, 3), so.contrast_loss_v2(x, 3, 50, kd), rtol=1e-6)


@pytest.mark.parametrize("name", ["bilinear_skip", "polynomial_16", "trilinear_A"])
def test_fusion_state_dict_layout_matches_reference(golden, name):
    """Fusion modules: same seed -> same initial weights as the reference constructor (global-RNG order: nn.Bilinear /
    nn.Linear defaults, then init_max_weights), same state_dict keys -- incl. the stage-2 PolynomialFusion."""
    g = golden(name)
    c = g.cfg
    kw = {k: v for k, v in c.items() if k not in ("B", "kind", "variant")}
    cls = {"bilinear": pkg.BilinearFusion, "polynomial": pkg.PolynomialFusion}.get(c["kind"]) or pkg.TrilinearFusion_A
    torch.manual_seed(2019)
    mod = cls(**kw)
    want, got = g.state_dict("init."), mod.state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert got[k].shape == want[k].shape, k
        if "encoder" not in k or ".0." in k:      # the goldens re-randomise the BatchNorm affine / running stats afterwards
            assert torch.equal(got[k], want[k]), k


def test_dropin_modules_resolve_to_the_package():
    import importlib
    import sys
    d = os.path.join(ROOT, "multimodal-learning_b200", "dropin")
    sys.path.insert(0, d)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in ("fusion", "KD_loss") or k.startswith("CL_utils")}
    try:
        fusion = importlib.import_module("fusion")
        crit = importlib.import_module("CL_utils.CRD_criterion")
        loss5 = importlib.import_module("CL_utils.CRD_loss")
        mem = importlib.import_module("CL_utils.memory_new")
        kd = importlib.import_module("KD_loss")
        assert fusion.BilinearFusion is pkg.BilinearFusion and fusion.PolynomialFusion is pkg.PolynomialFusion
        assert crit.CRDLoss is pkg.CRDLoss and loss5.CRDLoss is pkg.crd_select.CRDLoss
        assert mem.ContrastMemory_v3 is pkg.ContrastMemory_v3 and mem.ContrastMemory is pkg.ContrastMemory
        assert kd.DistillKL is pkg.DistillKL
    finally:
        sys.path.remove(d)
        for k in [k for k in sys.modules if k in ("fusion", "KD_loss") or k.startswith("CL_utils")]:
            sys.modules.pop(k)
        sys.modules.update(saved)


def test_graphed_train_step_rejects_what_it_cannot_capture():
    lin = torch.nn.Linear(4, 4)
    x = torch.zeros(2, 4)
    with pytest.raises(RuntimeError):        # CPU tensors: no CPU path
        pkg.GraphedTrainStep(lambda a: lin(a).sum(), lin.parameters(), torch.optim.SGD(lin.parameters(), lr=0.1), (x,))
