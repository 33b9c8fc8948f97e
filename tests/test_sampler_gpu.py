"""GPU tests of the device-side class-conditional index sampler (`mml_instance_sample`): the kernel's
integers equal the numpy restatement bit for bit, and the sampled tensor drives CRDLoss like the loader's `sample_idx`."""
from __future__ import annotations

import types

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import multimodal_learning_b200 as p
    assert torch.cuda.is_available()
    p._cabi.lib()
    return p


@pytest.mark.parametrize("n,C,B,mode,P,K", [
    (500, 3, 16, "exact", 1, 50),          # distinct negatives
    (500, 3, 16, "relax", 1, 700),         # with replacement (k > pool)
    (500, 3, 16, "multi_pos", 20, 100),
    (1024, 3, 16, "multi_pos", 300, 700),  # the reference's defaults (options.py:85-91)
    (100_000, 2, 8, "exact", 1, 16384),    # config-2-like width, wide Feistel domain
    (7, 2, 4, "exact", 1, 3),              # tiny pools
])
def test_kernel_equals_oracle_bit_for_bit(pkg, n, C, B, mode, P, K):
    rng = np.random.default_rng(n + K)
    labels = rng.integers(0, C, n)
    labels[:C] = np.arange(C)              # every class non-empty
    if mode == "multi_pos":
        B = min(B, n)
    idx = rng.permutation(n)[:B]
    seed = 0x1234_5678_9ABC_DEF0 & (2 ** 62 - 1)
    s = pkg.InstanceSampler(torch.from_numpy(labels), nce_k=K, nce_p=P, pos_mode=mode).cuda()
    got = s(torch.from_numpy(idx).to(DEV), seed=seed)
    want = so.instance_sample(idx, labels, C, P, K, mode, seed)
    assert got.dtype == torch.int64 and tuple(got.shape) == (B, P + K)
    assert np.array_equal(got.cpu().numpy(), want)
    # device seed word == host seed
    sd = torch.tensor([seed], dtype=torch.int64, device=DEV)
    from multimodal_learning_b200 import _cabi
    out2 = torch.empty_like(got)
    _cabi.check(_cabi.lib().mml_instance_sample(
        _cabi.dptr(torch.from_numpy(idx).to(DEV)), B, _cabi.dptr(s.labels), _cabi.dptr(s.order), _cabi.dptr(s.cls_ptr), C, n,
        s.p, K, {"exact": 0, "relax": 1, "multi_pos": 2}[mode], 0, _cabi.dptr(sd), _cabi.dptr(out2),
        _cabi.cur_stream(torch.device(DEV))), "mml_instance_sample")
    assert torch.equal(out2, got)


def test_survival_task_and_default_seed(pkg):
    n, K = 300, 200
    idx = torch.randperm(n)[:8]
    s = pkg.InstanceSampler(None, nce_k=K, task="surv", n_data=n)
    got = s(idx.to(DEV), seed=77)
    assert np.array_equal(got.cpu().numpy(), so.instance_sample(idx.numpy(), None, n, 1, K, "exact", 77))
    torch.manual_seed(3)
    a = s(idx.to(DEV))
    b = s(idx.to(DEV))
    torch.manual_seed(3)
    c = s(idx.to(DEV))
    assert not torch.equal(a, b) and torch.equal(a, c)           # fresh per call, reproducible under manual_seed


def test_negatives_are_uniform_over_the_pool(pkg):
    n, C, K, B = 64, 2, 40, 2000
    labels = np.arange(n) % C
    s = pkg.InstanceSampler(torch.from_numpy(labels), nce_k=K).cuda()
    idx = torch.zeros(B, dtype=torch.int64, device=DEV)          # same anchor (class 0): pool = the 32 odd samples
    torch.manual_seed(0)
    got = torch.cat([s(idx) for _ in range(5)])[:, 1:].cpu().numpy()
    assert (labels[got] == 1).all()
    cnt = np.bincount(got.reshape(-1), minlength=n)[1::2].astype(float)
    chi2 = ((cnt - cnt.mean()) ** 2 / cnt.mean()).sum()
    assert chi2 < 75                                              # 31 dof (K > pool: with replacement)


def test_sampled_indices_drive_crdloss(pkg):
    n, B, K = 4096, 32, 1024
    labels = torch.randint(0, 3, (n,))
    s = pkg.InstanceSampler(labels, nce_k=K).cuda()
    opt = types.SimpleNamespace(s_dim=32, t_dim=32, feat_dim=128, n_data=n, nce_k=K, nce_t=0.07, nce_m=0.5)
    crit = pkg.CRDLoss(opt).to(DEV)
    index = torch.randperm(n, device=DEV)[:B]
    sample_idx = s(index)
    assert torch.equal(sample_idx[:, 0], index)
    f_s = torch.randn(B, 32, device=DEV, requires_grad=True)
    loss = crit(f_s, torch.randn(B, 32, device=DEV), index, sample_idx)
    loss.backward()
    assert torch.isfinite(loss).all() and torch.isfinite(f_s.grad).all()
