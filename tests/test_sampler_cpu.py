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
