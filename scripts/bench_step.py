#!/usr/bin/env python
"""Fused distillation step (fusion -> CRD, forward + backward + bank update + Adam) at BASELINE configs 1 and 4 -- the
`fused_step` section of bench.py on its own.  One JSON line per config.

    python scripts/bench_step.py [--steps 20] [--no-cpu]"""
import argparse
import contextlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    for which, st in (("c1", a.steps * 10), ("c4", a.steps)):
        with contextlib.redirect_stdout(sys.stderr):
            rec = bench.fused_step_section(dev, which, st, a.warmup)
            if not a.no_cpu:
                torch.set_num_threads(bench.host_cores())
                rec["cpu_baseline"] = bench.cpu_fused_step("c1", 10, 2) if which == "c1" else bench.cpu_fused_step("c4", 2, 1, sample_B=64)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
