"""Host logic of the row-sharded bank under gloo, world_size 2, on CPU: all_gather / routing /
all_to_all / all_reduce / owner-update choreography with the oracle standing in for the kernels.
Every rank must reproduce the REFERENCE's single-bank golden results for the global batch."""
from __future__ import annotations

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,name", [(2, "crd_small"), (4, "crd_small"), (2, "crd_d128")])
def test_sharded_choreography_matches_reference_golden(tmp_path, world, name):
    out = tmp_path / "res.txt"
    port = 29600 + (os.getpid() + world * 7 + len(name)) % 300
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_sharded_worker.py"), str(world), "oracle", name,
                        str(port), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert out.read_text().startswith("ok")
