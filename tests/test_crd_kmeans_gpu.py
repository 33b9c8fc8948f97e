"""GPU parity tests of the per-class k-means centres (K13, csrc/crd_kmeans.cu; reference:
`MIA 2023/stage2_unimodal_student/CL_utils/CRD_criterion_v10.py:84-92, :122-129` = sklearn KMeans on the host).

The reference's fits start from sklearn's own random draws, so parity is: (1) from the initial centres recorded while the
unmodified reference ran (tests/golden/crdknn_kmeans_*), the CUDA path lands on sklearn's centres; (2) one Lloyd iteration
equals the numpy E/M step on ragged classes for every supported width; (3) properties at bank scale (fixed point, inertia
never increases, k-means++ draws are rows of the right class).  Tolerances: counts / assignments exact, floats 1e-5 abs on
unit-scale data (fp32 sums in a different but fixed order)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def km():
    import multimodal_learning_b200 as p
    assert torch.cuda.is_available()
    p._cabi.lib()
    return p.crd_kmeans


@pytest.fixture(scope="module")
def ko():
    from oracle import crd_knn_oracle
    return crd_knn_oracle


def _bank(n, D, sizes, seed, clustered=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D)).astype(np.float32) * 0.3
    if clustered:
        X += rng.standard_normal((clustered, D)).astype(np.float32)[rng.integers(0, clustered, n)]
    perm = rng.permutation(n)
    class_idx, at = [], 0
    for m in sizes:
        class_idx.append(np.sort(perm[at:at + m]))
        at += m
    return X, class_idx


def _numpy_step(X, class_idx, centres):
    new, counts, inertia, dist, sums = centres.copy(), [], [], [], np.zeros(centres.shape)
    for c, rows in enumerate(class_idx):
        x = X[rows].astype(np.float64)
        cen = centres[c].astype(np.float64)
        score = (cen * cen).sum(1)[None] - 2 * x @ cen.T
        lab = score.argmin(1)
        d = np.maximum((x * x).sum(1) + score.min(1), 0)
        dist.append(d)
        counts.append(np.bincount(lab, minlength=cen.shape[0]))
        inertia.append(np.bincount(lab, weights=d, minlength=cen.shape[0]))
        for j in range(cen.shape[0]):
            if (lab == j).any():
                new[c, j] = x[lab == j].mean(0)
                sums[c, j] = x[lab == j].sum(0)
    return new, np.stack(counts), np.stack(inertia), np.concatenate(dist), sums


@pytest.mark.parametrize("name", ["crdknn_kmeans_p4_d32", "crdknn_kmeans_p3_d128", "crdknn_kmeans_p6_d64"])
def test_kmeans_from_recorded_start_lands_on_sklearn_centres(km, golden, name):
    g = golden(name)
    c = g.cfg
    cls_of = g.np("row_class")
    class_idx = [np.nonzero(cls_of == k)[0] for k in range(3)]
    cls = km.ClassRows(class_idx, DEV)
    banks = {"memory_v1": g.t("init.contrast.memory_v1", DEV), "memory_v2": g.t("init.contrast.memory_v2", DEV)}
    for b, bank in enumerate(("memory_v1", "memory_v2")):       # step 0 fits the initial banks
        got, info = km.class_kmeans(banks[bank], cls, c["P"] - 1, init=g.t("step0.kmeans_init", DEV)[b], return_info=True)
        assert bool(info["done"].all())
        assert (got.cpu() - g.t("step0.kmeans_centres")[b]).abs().max() < 1e-5


@pytest.mark.parametrize("D,k,sizes", [
    (32, 1, [5, 1, 40]), (32, 8, [300, 17, 1000]), (64, 3, [129, 64, 33]), (128, 3, [1000, 777, 1]),
    (128, 7, [5000, 3000, 4100]), (256, 4, [900, 31, 650]), (512, 2, [260, 100, 7]), (128, 2, [16, 16]),
])
def test_one_lloyd_iteration_matches_numpy(km, D, k, sizes):
    n = sum(sizes) + 11
    X, class_idx = _bank(n, D, sizes, seed=D + k, clustered=k)
    rng = np.random.default_rng(1)
    init = np.stack([X[rng.choice(r, k, replace=len(r) < k)] for r in class_idx])
    want, counts, inertia, dist, sums = _numpy_step(X, class_idx, init)
    cls = km.ClassRows(class_idx, DEV)
    bank = torch.from_numpy(X).to(DEV)
    centres = torch.from_numpy(init).to(DEV).contiguous()
    C = len(sizes)
    got_in = torch.empty((C, k), device=DEV)
    got_n = torch.empty((C, k), dtype=torch.int64, device=DEV)
    got_d = torch.empty(sum(sizes), device=DEV)
    km.lloyd(bank, cls, centres, inertia=got_in, counts=got_n, row_dist=got_d)
    assert np.array_equal(got_n.cpu().numpy(), counts)
    assert np.abs(centres.cpu().numpy() - want).max() < 1e-5
    assert np.abs(got_d.cpu().numpy() - dist).max() < 1e-4 * max(1.0, dist.max())
    assert np.abs(got_in.cpu().numpy() - inertia).max() < 1e-4 * max(1.0, inertia.max())
    # assignment-only pass: centres stay, outputs describe them (same start again: the sums that shards would all-reduce)
    centres = torch.from_numpy(init).to(DEV).contiguous()
    got_s = torch.empty((C, k, D), device=DEV)
    km.lloyd(bank, cls, centres, update=False, counts=got_n, sums=got_s)
    assert torch.equal(torch.from_numpy(init).to(DEV), centres)
    assert np.array_equal(got_n.cpu().numpy(), counts)
    assert np.abs(got_s.cpu().numpy() - sums).max() < 1e-5 * max(1.0, np.abs(sums).max())


def test_kmeans_is_bit_reproducible_and_a_fixed_point(km):
    X, class_idx = _bank(60000, 128, [25000, 15000, 20000], seed=3, clustered=5)
    cls = km.ClassRows(class_idx, DEV)
    bank = torch.from_numpy(X).to(DEV)
    gen = torch.Generator(device=DEV)
    runs = []
    for _ in range(2):
        gen.manual_seed(7)
        runs.append(km.class_kmeans(bank, cls, 4, generator=gen, return_info=True))
    assert torch.equal(runs[0][0], runs[1][0])
    centres, info = runs[0]
    assert bool(info["done"].all()) and info["iterations_enqueued"] < 300
    again = centres.clone()
    inertia0 = torch.empty((3, 4), device=DEV)
    inertia1 = torch.empty((3, 4), device=DEV)
    km.lloyd(bank, cls, again, inertia=inertia0)                 # one more iteration moves no centre by more than the tolerance
    assert bool((((again - centres) ** 2).sum((1, 2)) <= info["tol"] * 10).all())
    km.lloyd(bank, cls, again, update=False, inertia=inertia1)
    assert bool((inertia1.sum(1) <= inertia0.sum(1) * (1 + 1e-5)).all())
    # and it beats its own start: inertia of the k-means++ centres >= inertia at the end
    gen.manual_seed(7)
    start = km.kmeans_plus_plus(bank, cls, 4, gen)
    inertia_s = torch.empty((3, 4), device=DEV)
    km.lloyd(bank, cls, start, update=False, inertia=inertia_s)
    assert bool((inertia1.sum(1) < inertia_s.sum(1)).all())


def test_kmeans_plus_plus_draws_rows_of_the_class(km):
    X, class_idx = _bank(5000, 64, [2000, 1200, 1800], seed=5, clustered=6)
    cls = km.ClassRows(class_idx, DEV)
    bank = torch.from_numpy(X).to(DEV)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(11)
    a = km.kmeans_plus_plus(bank, cls, 5, gen)
    gen.manual_seed(11)
    b = km.kmeans_plus_plus(bank, cls, 5, gen)
    assert torch.equal(a, b)
    a = a.cpu().numpy()
    for c, rows in enumerate(class_idx):
        for j in range(5):
            hit = np.nonzero((X[rows] == a[c, j]).all(1))[0]
            assert hit.size >= 1, (c, j)
        assert len({a[c, j].tobytes() for j in range(5)}) == 5          # D^2 sampling never re-draws a chosen row


def test_tolerance_is_sklearns(km, ko):
    X, class_idx = _bank(3000, 128, [1500, 400, 1100], seed=9, clustered=3)
    cls = km.ClassRows(class_idx, DEV)
    got = km.class_variance_tolerance(torch.from_numpy(X).to(DEV), cls).cpu().numpy()
    want = np.array([ko.kmeans_tolerance(X[r]) for r in class_idx])
    assert np.abs(got / want - 1).max() < 1e-4


def test_kmeans_rejects_what_it_cannot_do(km):
    import multimodal_learning_b200 as p
    X, class_idx = _bank(200, 48, [100, 100], seed=0)
    bank = torch.from_numpy(X).to(DEV)
    cls = km.ClassRows(class_idx, DEV)
    with pytest.raises(RuntimeError, match="feature width"):
        km.lloyd(bank, cls, torch.zeros((2, 2, 48), device=DEV))
    X, class_idx = _bank(200, 32, [100, 100], seed=0)
    bank = torch.from_numpy(X).to(DEV)
    cls = km.ClassRows(class_idx, DEV)
    with pytest.raises(NotImplementedError):
        km.lloyd(bank, cls, torch.zeros((2, 9, 32), device=DEV))
    with pytest.raises(RuntimeError):
        km.lloyd(bank.cpu(), cls, torch.zeros((2, 2, 32)))
    with pytest.raises(RuntimeError):
        km.ClassRows([[0, 1], []], DEV)
    bad = km.ClassRows([[0, 1, 2], [3, 4, 10 ** 6]], DEV)           # a row outside the bank: flagged, never read
    p._cabi.device_error_flags(reset=True)
    km.lloyd(bank, bad, torch.zeros((2, 2, 32), device=DEV))
    with pytest.raises(IndexError):
        p.check_device_errors()
