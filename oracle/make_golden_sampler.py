#!/usr/bin/env python
"""Generates tests/golden/instance_sampler.npz from the UNMODIFIED reference loader
`/root/reference/MICCAI-2022/data_loaders_MT.py` (`Pathomic_InstanceSample`, :146-256): the class-conditional pools its
constructor builds (:174-205) and the `sample_idx` rows its `__getitem__` draws (:222-249) for the three positive modes.
Test infrastructure only; run here (the reference is not on the GPU box), commit the fixture.

    python oracle/make_golden_sampler.py

Shims (nothing in the reference file is edited): a stub `utils` exposing `mixed_collate` (the real utils.py imports
lifelines / imblearn / torch_geometric, not installed); 16x16 PNGs written to a temp dir stand in for the pathology images
`__getitem__` opens.  numpy >= 1.24 refuses the ragged `np.asarray(list of arrays)` of :199-200, so the synthetic label
vector has equally sized classes (the pools are then rectangular); ragged class sizes are covered by the restatement test.
"""
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/MICCAI-2022"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "instance_sampler.npz")


def main():
    stub = types.ModuleType("utils")
    stub.mixed_collate = lambda batch: batch
    sys.modules["utils"] = stub
    sys.path.insert(0, REF)
    import data_loaders_MT as ref          # noqa: E402  (unmodified reference module)
    from PIL import Image

    n, ncls = 96, 3
    rng = np.random.default_rng(2019)
    labels = np.repeat(np.arange(ncls), n // ncls)
    rng.shuffle(labels)
    tmp = tempfile.mkdtemp()
    paths = []
    for i in range(n):
        p = os.path.join(tmp, f"{i}.png")
        Image.fromarray(rng.integers(0, 255, (16, 16, 3), dtype=np.uint8)).save(p)
        paths.append(p)
    data = {"train": {"x_path": np.asarray(paths), "x_omic": rng.normal(size=(n, 8)).astype(np.float32),
                      "e": np.zeros(n), "t": np.ones(n), "g": labels.astype(np.float64)}}
    out = {"labels": labels.astype(np.int64)}
    for mode, P, K in (("exact", 1, 40), ("relax", 1, 40), ("multi_pos", 6, 40), ("exact", 1, 200)):
        opt = types.SimpleNamespace(nce_p=P, nce_k=K, pos_mode=mode, task="grad", label_dim=3, input_size_path=8)
        ds = ref.Pathomic_InstanceSample(opt, data, "train")
        if "cls_positive" not in out:
            out["cls_positive"] = np.asarray(ds.cls_positive, dtype=np.int64)      # [3, 32]
            out["cls_negative"] = np.asarray(ds.cls_negative, dtype=np.int64)      # [3, 64]
        np.random.seed(7)
        idx = np.arange(0, n, 4)
        rows = np.stack([ds[int(i)][-1] for i in idx])
        tag = f"{mode}_P{P}_K{K}"
        out[tag + ".index"] = idx.astype(np.int64)
        out[tag + ".sample_idx"] = rows.astype(np.int64)
    # survival task (:222-227): every other sample is a candidate negative
    opt = types.SimpleNamespace(nce_p=1, nce_k=40, pos_mode="exact", task="surv", label_dim=3, input_size_path=8)
    ds = ref.Pathomic_InstanceSample(opt, data, "train")
    np.random.seed(11)
    idx = np.arange(0, n, 8)
    out["surv_K40.index"] = idx.astype(np.int64)
    out["surv_K40.sample_idx"] = np.stack([ds[int(i)][-1] for i in idx]).astype(np.int64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
