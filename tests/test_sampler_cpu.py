"""CPU checks of the class-conditional index sampler's restatement (oracle/sampler_oracle.py) against the reference's
contract (`MICCAI-2022/data_loaders_MT.py:174-205,222-249`): pools built exactly like the reference builds them, column 0 =
the anchor, positives / negatives from the right pools, distinct draws iff `k <= len(pool)`, uniform marginals."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import multimodal_learning_b200 as pkg
from oracle import sampler_oracle as so


@pytest.mark.parametrize("mode,P,K", [("exact", 1, 50), ("relax", 1, 700), ("multi_pos", 20, 100), ("multi_pos", 5, 700)])
def test_oracle_obeys_reference_pools(mode, P, K):
    rng = np.random.default_rng(0)
    labels = rng.integers(0, 3, 500)
    idx = rng.permutation(500)[:16]
    out = so.instance_sample(idx, labels, 3, P, K, mode, 1234567)
    pos, neg = so.reference_pools(labels, 3)
    order, ptr = so.class_tables(labels, 3)
    for c in range(3):      # the kernel's implicit pools ARE the reference's lists, in the reference's order
        assert np.array_equal(order[ptr[c]:ptr[c + 1]], pos[c])
        assert np.array_equal(np.concatenate((order[:ptr[c]], order[ptr[c + 1]:])), neg[c])
    for b in range(16):
        c = labels[idx[b]]
        if mode != "relax":
            assert out[b, 0] == idx[b]                                   # :229-230, :238
        assert np.isin(out[b, :P], pos[c]).all() and np.isin(out[b, P:], neg[c]).all()
        if K <= len(neg[c]):                                             # replace = k > len(pool), :243
            assert len(np.unique(out[b, P:])) == K
        if mode == "multi_pos":
            assert len(np.unique(out[b, 1:P])) == P - 1                  # replace=False, :237


def test_oracle_survival_task_excludes_the_anchor():
    rng = np.random.default_rng(1)
    idx = rng.permutation(400)[:8]
    o = so.instance_sample(idx, None, 400, 1, 300, "exact", 99)          # k <= n-1: distinct (:225-227)
    assert (o[:, 0] == idx).all()
    for b in range(8):
        assert len(np.unique(o[b, 1:])) == 300 and idx[b] not in o[b, 1:] and o[b, 1:].max() < 400
    o = so.instance_sample(idx, None, 400, 1, 900, "exact", 99)          # k > n-1: with replacement
    for b in range(8):
        assert idx[b] not in o[b, 1:] and o[b, 1:].max() < 400 and len(np.unique(o[b, 1:])) > 300


def test_keyed_bijection_is_a_permutation_with_uniform_marginals():
    for M in (1, 2, 3, 37, 64, 1000):
        key = [int(x) for x in so.philox4x32_10(0, 5, 7, 0, 11, 13)]
        assert sorted(so.perm_element(np.arange(M, dtype=np.uint64), M, key).tolist()) == list(range(M))
    M, cnt = 37, np.zeros(37)
    for s in range(3000):
        key = [int(x) for x in so.philox4x32_10(0, 5, s, 0, 1, 2)]
        cnt[so.perm_element(np.arange(3, dtype=np.uint64), M, key)] += 1
    chi2 = ((cnt - cnt.mean()) ** 2 / cnt.mean()).sum()
    assert chi2 < 70        # 36 dof: P(chi2 > 70) < 1e-3


def test_host_side_checks_mirror_numpy_errors():
    labels = torch.tensor([0, 0, 1, 1, 1, 2])
    with pytest.raises(ValueError):          # multi_pos asks for more distinct positives than the smallest class has (:237)
        pkg.InstanceSampler(labels, nce_k=4, nce_p=3, pos_mode="multi_pos")
    with pytest.raises(RuntimeError):        # one class only: empty negative pool
        pkg.InstanceSampler(torch.zeros(5, dtype=torch.long), nce_k=4)
    with pytest.raises(NotImplementedError):
        pkg.InstanceSampler(labels, nce_k=4, pos_mode="nearest")
    s = pkg.InstanceSampler(labels, nce_k=4)
    with pytest.raises(RuntimeError):        # no CPU fallback
        s(torch.tensor([0, 1]))


# ---------------------------------------------------------------------------------------------------------------- #
# pinned against the reference loader itself: tests/golden/instance_sampler.npz, written by oracle/make_golden_sampler.py
# from the UNMODIFIED MICCAI-2022/data_loaders_MT.py (Pathomic_InstanceSample.__init__ / __getitem__)
# ---------------------------------------------------------------------------------------------------------------- #
def _contract(rows, index, labels, pos_pools, neg_pools, mode, P, K):
    """The reference's sample_idx contract (:229-249), as a predicate on ANY sampler's output."""
    for b in range(len(index)):
        c = labels[index[b]]
        if mode != "relax":
            assert rows[b, 0] == index[b]
        assert np.isin(rows[b, :P], pos_pools[c]).all()
        assert np.isin(rows[b, P:], neg_pools[c]).all()
        if K <= len(neg_pools[c]):
            assert len(np.unique(rows[b, P:])) == K                      # np.random.choice(..., replace=False)
        if mode == "multi_pos":
            assert len(np.unique(rows[b, 1:P])) == P - 1


def test_class_tables_equal_the_reference_loaders_pools():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "instance_sampler.npz"))
    labels = g["labels"]
    order, ptr = so.class_tables(labels, 3)
    pos, neg = so.reference_pools(labels, 3)
    for c in range(3):
        # the kernel's implicit pools (a segment of `order`; `order` with that segment cut out) ARE the lists the reference's
        # constructor built, element for element and in the same order
        assert np.array_equal(order[ptr[c]:ptr[c + 1]], g["cls_positive"][c])
        assert np.array_equal(np.concatenate((order[:ptr[c]], order[ptr[c + 1]:])), g["cls_negative"][c])
        assert np.array_equal(pos[c], g["cls_positive"][c]) and np.array_equal(neg[c], g["cls_negative"][c])


@pytest.mark.parametrize("mode,P,K", [("exact", 1, 40), ("relax", 1, 40), ("multi_pos", 6, 40), ("exact", 1, 200)])
def test_reference_draws_and_ours_obey_the_same_contract(mode, P, K):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "instance_sampler.npz"))
    labels, tag = g["labels"], f"{mode}_P{P}_K{K}"
    index, ref_rows = g[tag + ".index"], g[tag + ".sample_idx"]
    assert ref_rows.shape == (len(index), P + K)
    pos, neg = list(g["cls_positive"]), list(g["cls_negative"])
    _contract(ref_rows, index, labels, pos, neg, mode, P, K)             # what the reference's numpy draws satisfy ...
    ours = so.instance_sample(index, labels, 3, P, K, mode, 20221231)
    _contract(ours, index, labels, pos, neg, mode, P, K)                 # ... the device sampler's algorithm satisfies too
    # same marginal law: pooled over anchors, negatives are uniform over the other classes in both samplers
    if K <= 64:
        for rows in (ref_rows, ours):
            cnt = np.bincount(rows[:, P:].ravel(), minlength=96).astype(float)
            assert cnt.max() <= 3 * cnt[cnt > 0].mean() + 8


def test_survival_task_contract_on_reference_draws():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "instance_sampler.npz"))
    index, rows = g["surv_K40.index"], g["surv_K40.sample_idx"]
    ours = so.instance_sample(index, None, 96, 1, 40, "exact", 5)
    for r in (rows, ours):
        assert (r[:, 0] == index).all()
        for b in range(len(index)):
            assert index[b] not in r[b, 1:] and len(np.unique(r[b, 1:])) == 40 and r[b, 1:].max() < 96


def test_keyed_bijection_pairs_are_jointly_uniform():
    """np.random.choice(replace=False) is uniform over ORDERED K-subsets; marginal uniformity of each position is not
    enough.  First two images of the keyed bijection of [0, 12): all 132 ordered pairs equally likely (chi-square over
    26400 keys, 131 degrees of freedom), and the third image is uniform given the first."""
    M, n_keys = 12, 26400
    cnt = np.zeros((M, M))
    third = np.zeros((M, M))
    for s in range(n_keys):
        key = [int(x) for x in so.philox4x32_10(0, 5, s & 0xFFFF, s >> 16, 17, 23)]
        a, b, c = so.perm_element(np.arange(3, dtype=np.uint64), M, key)
        cnt[a, b] += 1
        third[a, c] += 1
    assert np.trace(cnt) == 0 and np.trace(third) == 0                   # a bijection never repeats an element
    off = ~np.eye(M, dtype=bool)
    for table in (cnt, third):
        e = n_keys / (M * (M - 1))
        chi2 = ((table[off] - e) ** 2 / e).sum()
        assert chi2 < 190                                                # 131 dof: P(chi2 > 190) < 6e-4
