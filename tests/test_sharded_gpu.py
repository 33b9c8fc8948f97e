"""Row-sharded bank on real GPUs (NCCL, one process per GPU): sharded result == reference golden of the
single-bank module.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box."""
from __future__ import annotations

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("transport", ["peer", "alltoall"])
@pytest.mark.parametrize("name", ["crd_small", "crd_d128"])
def test_sharded_cuda_matches_reference_golden(tmp_path, name, transport):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    while {"crd_small": 8, "crd_d128": 16}[name] % world:
        world -= 1
    out = tmp_path / "res.txt"
    port = 29500 + os.getpid() % 400
    env = dict(os.environ, MML_TRANSPORT=transport)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_sharded_worker.py"), str(world), "cuda", name,
                        str(port), str(out)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = out.read_text()
    assert res.startswith("ok") and f"transport={transport}" in res, res + r.stdout[-1500:]


def test_sharded_step_replayed_from_cuda_graph_equals_eager(tmp_path):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    out = tmp_path / "res.txt"
    port = 29900 + os.getpid() % 90
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_sharded_graph_worker.py"), str(world), str(port),
                        str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert out.read_text().startswith("ok")


def test_sharded_knn_positives_equal_single_bank(tmp_path):
    """N4 over the sharded bank: per-shard exact top-P + candidate exchange == `knn_positives` over the whole bank."""
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    out = tmp_path / "res.txt"
    port = 29800 + os.getpid() % 90
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_sharded_knn_worker.py"), str(world), str(port), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert out.read_text().startswith("ok")


def test_sharded_class_kmeans_equals_single_bank(tmp_path):
    """N4 "centers" over the sharded bank: per-shard assign pass + all-reduced sums == `class_kmeans` over the whole bank."""
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    out = tmp_path / "res.txt"
    port = 29700 + os.getpid() % 90
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_sharded_kmeans_worker.py"), str(world), str(port), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert out.read_text().startswith("ok")
