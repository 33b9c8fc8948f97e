"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the device-side class-conditional index sampler
(`multimodal-learning_b200/csrc/sampler.cu`) and of the pools it draws from
(`MICCAI-2022/data_loaders_MT.py:174-205` cls_positive / cls_negative, `:222-249` the draws).

Parity status: the reference draws with numpy's mt19937 inside DataLoader workers; that stream cannot be reproduced by
a counter-based GPU generator, so **the RNG stream is unpinned by construction**.  What IS pinned:
  * integer work, bit for bit: the kernel's output equals `instance_sample` below for the same seed
    (Philox4x32-10 + multiply-shift for draws with replacement; keyed 8-round Feistel bijection + cycle walking for
    draws without replacement);
  * the reference's contract, checked against pools built exactly like the reference builds them
    (`reference_pools`): column 0 = the anchor, positives from the anchor's class, negatives from the other classes,
    distinct iff `k <= len(pool)` (`replace = k > len(pool)`, :243), uniform marginals.
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all arguments broadcastable uint32 arrays.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & M32 for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = np.uint64(int(k0) & 0xFFFFFFFF), np.uint64(int(k1) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & M32
        n1 = p1 & M32
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & M32
        n3 = p0 & M32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return c0, c1, c2, c3


def mix32(h):
    h = np.asarray(h, dtype=np.uint64) & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x7feb352d)) & M32
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x846ca68b)) & M32
    h ^= h >> np.uint64(16)
    return h


def perm_element(j, M, key):
    """j-th element (j array) of the keyed bijection of [0, M); key = 4 uint32 scalars."""
    M = int(M)
    bits = max(int(M - 1).bit_length(), 1)
    if M <= 2:
        bits = 2
    bits = (bits + 1) & ~1
    half = np.uint64(bits >> 1)
    mask = np.uint64((1 << int(half)) - 1)
    x = np.asarray(j, dtype=np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        L, R = x[todo] >> half, x[todo] & mask
        for r in range(8):      # 8 rounds: 4 leave the first two images of a small domain far from jointly uniform
            rk = np.uint64((int(key[r & 3]) + 0x9E3779B9 * (r >> 2)) & 0xFFFFFFFF)
            t = L ^ (mix32(R ^ rk) & mask)
            L, R = R, t
        x[todo] = (L << half) | R
        todo = x >= np.uint64(M)
    return x.astype(np.int64)


def bounded(r, M):
    return ((np.asarray(r, dtype=np.uint64) * np.uint64(int(M))) >> np.uint64(32)).astype(np.int64)


def class_tables(labels, num_classes):
    """order (sample ids sorted by class, stable) and cls_ptr -- the kernel's view of the reference's pools."""
    labels = np.asarray(labels)
    order = np.argsort(labels, kind="stable").astype(np.int32)
    cls_ptr = np.zeros(num_classes + 1, dtype=np.int32)
    cls_ptr[1:] = np.cumsum(np.bincount(labels, minlength=num_classes))
    return order, cls_ptr


def reference_pools(labels, num_classes):
    """cls_positive / cls_negative exactly as data_loaders_MT.py:193-202 builds them."""
    n = len(labels)
    cls_positive = [[] for _ in range(num_classes)]
    for i in range(n):
        cls_positive[labels[i]].append(i)
    cls_negative = [[] for _ in range(num_classes)]
    for i in range(num_classes):
        for j in range(num_classes):
            if j == i:
                continue
            cls_negative[i].extend(cls_positive[j])
    return [np.asarray(p) for p in cls_positive], [np.asarray(p) for p in cls_negative]


def instance_sample(index, labels, num_classes, P, K, pos_mode, seed):
    """The kernel's algorithm in numpy: -> int64 [B, P+K].  labels=None is the survival task."""
    index = np.asarray(index, dtype=np.int64)
    B = len(index)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    out = np.zeros((B, P + K), dtype=np.int64)
    mode = {"exact": 0, "relax": 1, "multi_pos": 2}[pos_mode]
    if labels is not None:
        order, cls_ptr = class_tables(labels, num_classes)
        n = len(labels)
    for b in range(B):
        anchor = int(index[b])
        blo, bhi = b & 0xFFFFFFFF, (b >> 32) & 0xFFFFFFFF
        jn = np.arange(K, dtype=np.uint64)
        if labels is None:
            n_all = int(num_classes)            # surv task: `num_classes` carries n
            M = n_all - 1
            out[b, :P] = anchor
            if K > M:
                e = bounded(philox4x32_10(jn, 0, blo, bhi ^ 0x10000000, k0, k1)[0], M)
            else:
                key = [int(x) for x in philox4x32_10(0, 1, blo, bhi ^ 0x20000000, k0, k1)]
                e = perm_element(jn, M, key)
            out[b, P:] = np.where(e < anchor, e, e + 1)
            continue
        c = int(labels[anchor])
        seg0, seg1 = int(cls_ptr[c]), int(cls_ptr[c + 1])
        Mp, Mn = seg1 - seg0, n - (seg1 - seg0)
        if mode == 0:
            out[b, :P] = anchor
        elif mode == 1:
            r = philox4x32_10(np.arange(P, dtype=np.uint64), 2, blo, bhi ^ 0x30000000, k0, k1)[0]
            out[b, :P] = order[seg0 + bounded(r, Mp)]
        else:
            key = [int(x) for x in philox4x32_10(0, 3, blo, bhi ^ 0x40000000, k0, k1)]
            out[b, :P] = order[seg0 + perm_element(np.arange(P, dtype=np.uint64), Mp, key)]
            out[b, 0] = anchor
        if K > Mn:
            e = bounded(philox4x32_10(jn, 4, blo, bhi ^ 0x50000000, k0, k1)[0], Mn)
        else:
            key = [int(x) for x in philox4x32_10(0, 5, blo, bhi ^ 0x60000000, k0, k1)]
            e = perm_element(jn, Mn, key)
        out[b, P:] = order[np.where(e < seg0, e, e + Mp)]
    return out
