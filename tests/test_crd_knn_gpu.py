"""GPU parity tests of the KNN / class-centre CRD criterion (reference: `MIA 2023/stage2_unimodal_student/CL_utils/
CRD_criterion_v10.py`): the CUDA path (through the C ABI and the drop-in modules) vs the reference-generated goldens
(oracle/make_golden_knn.py) and the CPU oracle.

Tolerances: neighbour rows bit-exact; floats rel 1e-4 (fp32 memory path; the TF32 pass only nominates candidates, every
reported similarity is recomputed in fp32)."""
from __future__ import annotations

import types

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"
CASES = ["crdknn_p3_d16", "crdknn_p5_d128", "crdknn_p1_d32", "crdknn_centers_d32", "crdknn_kmeans_p4_d32", "crdknn_kmeans_p3_d128",
         "crdknn_kmeans_p6_d64"]


@pytest.fixture(scope="module")
def pkg():
    import multimodal_learning_b200 as p
    assert torch.cuda.is_available()
    p._cabi.lib()
    return p


@pytest.fixture(scope="module")
def ko():
    from oracle import crd_knn_oracle
    return crd_knn_oracle


@pytest.mark.parametrize("name", CASES)
def test_crd_knn_module_matches_reference_golden(pkg, golden, name, capsys):
    g = golden(name)
    c = g.cfg
    cls = g.np("row_class")
    class_idx = [np.nonzero(cls == k)[0] for k in range(3)]
    opt = types.SimpleNamespace(s_dim=c["s_dim"], t_dim=c["t_dim"], feat_dim=c["D"], nce_k=c["K"], nce_t=c["T"],
                                nce_m=c["momentum"], nce_p=c["P"], pos_extra=c["pos_extra"])
    mod = pkg.crd_knn.CRDLoss(opt, c["n"], class_idx)
    assert torch.equal(mod.contrast.all_sample_labels, torch.as_tensor(cls).float())
    mod.load_state_dict(g.state_dict("init."))
    mod = mod.to(DEV)
    captured = {}
    mod.contrast.register_forward_hook(lambda m, i, o: captured.update(out=o))
    before = pkg._cabi.launch_count()
    for s in range(c["steps"]):
        p = f"step{s}."
        f_s = g.t(p + "f_s", DEV).requires_grad_(True)
        f_t = g.t(p + "f_t", DEV).requires_grad_(True)
        pre1 = mod.contrast.memory_v1.clone()
        mod.zero_grad()
        if c["pos_extra"] == "centers" and c["P"] > 2:      # the reference's k-means starts at random: same start as recorded
            mod.contrast.kmeans_init = g.t(p + "kmeans_init", DEV)
        loss, sample_loss = mod(g.t(p + "sample_weights", DEV), f_s, f_t, g.t(p + "label", DEV), g.t(p + "idx", DEV),
                                g.t(p + "contrast_idx", DEV))
        loss.backward()
        out = captured["out"]
        assert out[0].is_contiguous() and out[0].shape == g.t(p + "out_v1").shape
        assert rel_err(out[0], g.t(p + "out_v1")) < TOL and rel_err(out[1], g.t(p + "out_v2")) < TOL
        if c["pos_extra"] == "neighbors":
            assert rel_err(out[2], g.t(p + "sim_v1")) < TOL and rel_err(out[3], g.t(p + "sim_v2")) < TOL
        assert rel_err(loss.reshape(-1), g.t(p + "loss")) < TOL
        assert rel_err(sample_loss, g.t(p + "sample_loss")) < TOL
        assert rel_err(f_s.grad, g.t(p + "grad_f_s")) < TOL
        assert rel_err(f_t.grad, g.t(p + "grad_f_t")) < TOL
        for k, v in mod.named_parameters():
            assert rel_err(v.grad, g.t(p + "grad." + k)) < TOL, k
        assert rel_err(mod.contrast.params, g.t(p + "params")) < TOL
        assert rel_err(mod.contrast.memory_v1, g.t(p + "memory_v1")) < TOL
        assert rel_err(mod.contrast.memory_v2, g.t(p + "memory_v2")) < TOL
        changed = (mod.contrast.memory_v1 != pre1).any(dim=1).nonzero().flatten().cpu().tolist()
        assert sorted(changed) == sorted(g.t(p + "idx").tolist())
    assert pkg._cabi.launch_count() > before
    assert "normalization constant Z_v1 is set to" in capsys.readouterr().out


def _problem(n, D, B, n_cls=3, seed=0, clustered=False):
    gen = torch.Generator().manual_seed(seed)
    bank = torch.randn(n, D, generator=gen)
    if clustered:      # many near-duplicates: tight clusters around a few centres, the worst case for a TF32 filter
        centres = torch.randn(12, D, generator=gen)
        bank = centres[torch.randint(0, 12, (n,), generator=gen)] + 1e-3 * bank
    bank = bank * (0.5 + torch.rand(n, 1, generator=gen))            # rows are not unit-norm before their first update
    labels = torch.randint(0, n_cls, (n,), generator=gen)
    rows = torch.randperm(n, generator=gen)[:B]
    return bank, labels, rows, labels[rows].clone()


@pytest.mark.parametrize("n,D,B,P,clustered", [
    (20000, 128, 300, 5, False),      # several anchor tiles, many bank slices
    (1000, 128, 16, 3, False),        # the reference's real scale: one slice
    (777, 64, 5, 8, False),           # ragged last tile, P = the maximum
    (5000, 96, 130, 1, False),        # P = 1: the query itself
    (3000, 32, 40, 4, False),
    (2500, 48, 33, 4, False),         # D not a multiple of 32: exact scan only
    (6000, 128, 64, 5, True),         # clustered bank: the error bound cannot prove the filter -> exact scan for those anchors
])
def test_knn_kernel_matches_oracle_bit_for_bit(pkg, ko, n, D, B, P, clustered):
    bank, labels, rows, blab = _problem(n, D, B, clustered=clustered)
    more_idx, more_sim = ko.knn_neighbors(bank.double(), rows, labels, blab, min(P + 1, n))      # float64: the exact ranking
    want_idx, want_sim = more_idx[:, :P], more_sim[:, :P]
    bd, ld = bank.to(DEV), labels.to(DEV, torch.int32)
    for exact_only, ncls in ((False, 3), (False, 0), (True, 0)):        # tabulated class mask / label compares / exact scan
        idx, sim, flags = pkg.crd_knn.knn_positives(bd, ld, rows.to(DEV), blab.to(DEV), P, n_classes=ncls,
                                                    exact_only=exact_only, return_flags=True)
        # neighbours whose float64 similarities differ by less than fp32 resolution may swap: compare through the values
        got = ko.masked_cosine(bank.double(), rows, labels, blab).gather(1, idx.cpu())
        assert (got - want_sim).abs().max() < 5e-6, (exact_only, (got - want_sim).abs().max())
        assert (sim.cpu().double() - got).abs().max() < 5e-6
        clear = (more_sim[:, :-1] - more_sim[:, 1:]).min(dim=1).values > 1e-5      # every gap down to the first row left out
        assert torch.equal(idx.cpu()[clear], want_idx[clear])                        # unambiguous rankings: bit-exact rows
        if exact_only or D % 32 != 0:
            assert int(flags.sum()) == B
        elif clustered:
            assert int(flags.sum()) > 0
        else:
            assert int(flags.sum()) <= B // 10
    assert (idx[:, 0].cpu() == rows).all() or clustered                              # the best neighbour is the query itself
    # caller-kept inverse norms (what the module does between steps): same rows, same similarities
    inv = pkg.crd_knn.knn_inv_norms(bd)
    assert rel_err(inv, 1.0 / bank.norm(dim=1)) < 1e-6
    part = pkg.crd_knn.knn_inv_norms(bd, torch.zeros(n, device=DEV), rows=rows.to(DEV))
    assert torch.equal(part[rows.to(DEV)], inv[rows.to(DEV)]) and int((part != 0).sum()) == B
    idx2, sim2 = pkg.crd_knn.knn_positives(bd, ld, rows.to(DEV), blab.to(DEV), P, n_classes=3, inv_norms=inv)
    idx3, sim3 = pkg.crd_knn.knn_positives(bd, ld, rows.to(DEV), blab.to(DEV), P, n_classes=3)
    assert torch.equal(idx2, idx3) and torch.equal(sim2, sim3)
    # explicit query vectors (the sharded bank's case) == the bank's own rows
    idx4, sim4 = pkg.crd_knn.knn_positives(bd, ld, None, blab.to(DEV), P, n_classes=3, queries=bd[rows.to(DEV)].contiguous())
    assert torch.equal(idx4, idx3) and torch.equal(sim4, sim3)


def test_knn_fewer_same_class_rows_than_positives(pkg, ko):
    """Other classes score exactly 0 (the reference multiplies by a 0/1 mask): with only 2 rows of the anchor's class and
    negative cosines around, the zeros of the other classes are next -- smallest row first."""
    n, D, P = 400, 32, 4
    gen = torch.Generator().manual_seed(3)
    bank = torch.randn(n, D, generator=gen)
    labels = torch.ones(n, dtype=torch.long)
    labels[[17, 250]] = 0
    rows, blab = torch.tensor([17, 250]), torch.tensor([0, 0])
    want_idx, want_sim = ko.knn_neighbors(bank, rows, labels, blab, P)
    for ncls in (2, 0):
        idx, sim = pkg.crd_knn.knn_positives(bank.to(DEV), labels.to(DEV, torch.int32), rows.to(DEV), blab.to(DEV), P, n_classes=ncls)
        assert torch.equal(idx.cpu(), want_idx) and (sim.cpu() - want_sim).abs().max() < 1e-6
    assert torch.equal(idx.cpu(), want_idx) and (sim.cpu() - want_sim).abs().max() < 1e-6
    assert idx[0, 0] == 17 and (sim[:, 2:] == 0).any()


def test_knn_rejects_what_it_cannot_do(pkg):
    bank = torch.randn(100, 32, device=DEV)
    lab = torch.zeros(100, dtype=torch.int32, device=DEV)
    rows = torch.arange(4, device=DEV)
    with pytest.raises(NotImplementedError):
        pkg.crd_knn.knn_positives(bank, lab, rows, rows * 0, 9)
    with pytest.raises(RuntimeError):
        pkg.crd_knn.knn_positives(bank.cpu(), lab, rows, rows * 0, 3)
    pkg.crd_knn.knn_positives(bank, lab, torch.tensor([0, 1, 2, 100], device=DEV), rows * 0, 3)     # row 100 does not exist
    assert pkg.device_error_flags(reset=True) & pkg._cabi.DEVERR_CRD_INDEX


def test_knn_at_config2_scale_sampled_vs_exact(pkg):
    """n = 1M rows x 128, 1024 anchors (BASELINE config 2's bank): a sample of anchors against an exact fp32 top-k."""
    n, D, B, P = 1_000_000, 128, 1024, 5
    gen = torch.Generator(device=DEV).manual_seed(1)
    bank = torch.randn(n, D, device=DEV, generator=gen)
    bank = bank / bank.norm(dim=1, keepdim=True)
    labels = torch.randint(0, 3, (n,), device=DEV, generator=gen, dtype=torch.int32)
    rows = torch.randperm(n, device=DEV, generator=gen)[:B]
    blab = labels[rows].long()
    idx, sim, flags = pkg.crd_knn.knn_positives(bank, labels, rows, blab, P, n_classes=3, return_flags=True)
    assert int(flags.sum()) <= 8
    pick = torch.arange(0, B, 37, device=DEV)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        full = bank[rows[pick]] @ bank.t()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    full = full * (labels.view(1, -1) == blab[pick].view(-1, 1)).float()
    want_sim, want_idx = full.topk(P, dim=1)
    assert (sim[pick] - want_sim).abs().max() < 5e-6
    assert (full.gather(1, idx[pick]) - want_sim).abs().max() < 5e-6
    assert (idx[pick, 0] == rows[pick]).all()
