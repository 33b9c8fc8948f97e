"""K13 per-class k-means centres at bank scale: one Lloyd iteration (HBM pass over the bank) and the whole fit, against
sklearn on the host for one class (the reference's path, CRD_criterion_v10.py:84-92).  One JSON line per case."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_learning_b200 as pkg  # noqa: E402
from multimodal_learning_b200 import crd_kmeans as km  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
    peak = float(peaks.get("hbm_gbs", 0) or 0) or None
    cpu = "--cpu" in sys.argv
    for n, D, k in ((1 << 20, 128, 3), (1 << 20, 128, 7), (1 << 18, 128, 3)):
        g = torch.Generator(device=dev)
        g.manual_seed(0)
        modes = torch.randn(24, D, device=dev, generator=g)
        bank = torch.nn.functional.normalize(modes[torch.randint(0, 24, (n,), device=dev, generator=g)]
                                             + 0.5 * torch.randn(n, D, device=dev, generator=g), dim=1).contiguous()
        labels = torch.randint(0, 3, (n,), device=dev, generator=g)
        class_idx = [torch.nonzero(labels == c).flatten().cpu().numpy() for c in range(3)]
        cls = km.ClassRows(class_idx, dev)
        g.manual_seed(1)
        start = km.kmeans_plus_plus(bank, cls, k, g)
        centres = start.clone()
        ws = km.lloyd(bank, cls, centres)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        km.lloyd(bank, cls, centres, iterations=reps, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        it_ms = e0.elapsed_time(e1) / reps
        bytes_it = n * D * 4 + n * 8
        g.manual_seed(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fit, info = km.class_kmeans(bank, cls, k, generator=g, return_info=True)
        torch.cuda.synchronize()
        fit_ms = (time.perf_counter() - t0) * 1e3
        line = {"op": "class_kmeans", "n": n, "D": D, "classes": 3, "k": k, "lloyd_iteration_ms": round(it_ms, 4),
                "GBps": round(bytes_it / it_ms / 1e6, 1), "frac_of_hbm_peak": round(bytes_it / it_ms / 1e6 / peak, 3) if peak else None,
                "fit_ms": round(fit_ms, 2), "iterations_enqueued": info["iterations_enqueued"], "converged": bool(info["done"].all())}
        if cpu and n == 1 << 18:
            from sklearn.cluster import KMeans
            X = bank[torch.as_tensor(class_idx[0], device=dev)].cpu().numpy()
            t0 = time.perf_counter()
            est = KMeans(n_clusters=k).fit(X)
            line["cpu_sklearn_one_class_s"] = round(time.perf_counter() - t0, 2)
            line["cpu_sklearn_rows"] = int(X.shape[0])
            line["cpu_sklearn_iters"] = int(est.n_iter_)
        print(json.dumps(line), flush=True)
    pkg.check_device_errors()


if __name__ == "__main__":
    main()
