"""Shared fixtures.  `-m "not gpu"` runs here (no GPU); `-m gpu` runs on a B200."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


class Golden:
    """One tests/golden/*.npz fixture: `.cfg` (dict) + tensors by key."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.cfg = json.loads(bytes(z["__config__"]).decode())
        self._z = z

    def keys(self):
        return [k for k in self._z.files if k != "__config__"]

    def np(self, key):
        return self._z[key]

    def t(self, key, device="cpu"):
        return torch.from_numpy(self._z[key].copy()).to(device)

    def state_dict(self, prefix="init.", device="cpu"):
        return {k[len(prefix):]: self.t(k, device) for k in self.keys() if k.startswith(prefix)}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]
    return load


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| -- the 'rel' every tolerance in this suite refers to."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    denom = b.norm().item()
    return (a - b).norm().item() / (denom if denom > 0 else 1.0)
