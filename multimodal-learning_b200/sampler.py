"""Class-conditional contrast indices on the device.

The reference produces `index, sample_idx` per sample inside `Pathomic_InstanceSample.__getitem__`
(`MICCAI-2022/data_loaders_MT.py:222-256`: numpy draws over O(n) candidate lists in DataLoader workers) and uploads the
collated `[B, nce_p + nce_k]` int64 tensor every step (`train_test_MT.py:164`).  `InstanceSampler` keeps the labels on the
GPU and writes the same tensor with one kernel (`mml_instance_sample`):

    sampler = InstanceSampler(labels, nce_k=opt.nce_k, nce_p=opt.nce_p, pos_mode=opt.pos_mode, task=opt.task).cuda()
    sample_idx = sampler(index.cuda())                 # [B, nce_p + nce_k] int64, column 0 = the anchor
    loss = criterion(f_s, f_t, index.cuda(), sample_idx)

Same pools and rules as the reference (class-conditional positives / negatives, `replace = k > len(pool)`, `pos_idx[0] =
index`); the random stream is a counter-based Philox keyed by a device-side seed word (reproducible under
`torch.manual_seed`, CUDA-graph safe), not numpy's mt19937 -- see oracle/sampler_oracle.py for what is pinned.
"""
from __future__ import annotations

import torch

from . import _cabi

_POS_MODES = {"exact": 0, "relax": 1, "multi_pos": 2}


class InstanceSampler:
    def __init__(self, labels, nce_k, nce_p=1, pos_mode="exact", task="grad", num_classes=None, n_data=None):
        if pos_mode not in _POS_MODES:
            raise NotImplementedError(pos_mode)                      # data_loaders_MT.py:240
        self.k, self.pos_mode, self.task = int(nce_k), pos_mode, task
        self.p = int(nce_p) if pos_mode == "multi_pos" else 1        # 'exact' / 'relax' yield one positive (:229-233)
        if task == "surv":
            self.n = int(n_data if n_data is not None else len(labels))
            self.labels = self.order = self.cls_ptr = None
            self.num_classes = 0
            if pos_mode != "exact":
                raise RuntimeError("the survival task uses pos_idx = index (data_loaders_MT.py:223)")
        else:
            lab = torch.as_tensor(labels).to(torch.int64).cpu()
            self.n = lab.numel()
            self.num_classes = int(num_classes if num_classes is not None else int(lab.max().item()) + 1)
            counts = torch.bincount(lab, minlength=self.num_classes)
            if (counts == self.n).any():
                raise RuntimeError("every sample is in one class: cls_negative is empty (np.random.choice would raise)")
            if pos_mode == "multi_pos" and int(counts[counts > 0].min().item()) < self.p:
                raise ValueError("Cannot take a larger sample than population when 'replace=False'")   # numpy's message (:237)
            self.labels = lab.to(torch.int32)
            self.order = torch.argsort(lab, stable=True).to(torch.int32)
            ptr = torch.zeros(self.num_classes + 1, dtype=torch.int32)
            ptr[1:] = counts.cumsum(0)
            self.cls_ptr = ptr
        if self.n >= 2 ** 31:
            raise RuntimeError("n_data must fit in int32")

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device()) if device is None else device)

    def to(self, device):
        if self.labels is not None:
            self.labels, self.order, self.cls_ptr = (t.to(device) for t in (self.labels, self.order, self.cls_ptr))
        return self

    def __call__(self, index, seed=None):
        """index: int64 [B] CUDA tensor -> sample_idx int64 [B, p + k].  `seed` (int) pins the draw for tests; by default
        one 64-bit word per call is drawn on the device from torch's CUDA generator."""
        if not index.is_cuda:
            raise RuntimeError("InstanceSampler runs on CUDA tensors only (no CPU fallback)")
        if self.labels is not None and not self.labels.is_cuda:
            raise RuntimeError("move the sampler to the GPU first: InstanceSampler(...).cuda()")
        index = index.to(torch.int64).contiguous()
        B = index.numel()
        out = torch.empty(B, self.p + self.k, dtype=torch.int64, device=index.device)
        seed_t = None
        if seed is None:
            seed_t = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=index.device)
            seed = 0
        rc = _cabi.lib().mml_instance_sample(
            _cabi.dptr(index), B, _cabi.dptr(self.labels), _cabi.dptr(self.order), _cabi.dptr(self.cls_ptr),
            self.num_classes, self.n, self.p, self.k, _POS_MODES[self.pos_mode], int(seed), _cabi.dptr(seed_t),
            _cabi.dptr(out), _cabi.cur_stream(index.device))
        _cabi.check(rc, "mml_instance_sample")
        return out
