"""Benchmark of the multimodal distillation hot path (BASELINE.json metric: distill steps/s; % HBM / tensor roofline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path on the host cores

ONE JSON line on stdout (everything else goes to stderr).  Headline workload (`config.workload`, `value`, `e2e`):
  N=1  BASELINE config 2: feat_dim 128, batch 1024, n_data 1M, nce_k 16384 -- one step = Embed(f_s), Embed(f_t) ->
       fused gather / score / NCE loss / gradient over 2 x 1024 x 16385 bank rows -> momentum row update ->
       loss.backward() through the Embed heads (f_t detached) -> Adam step on the Embed parameters.
  N>1  the same unit of work per GPU on the row-sharded bank (2M rows per GPU; N=8 is exactly BASELINE config 5:
       16M x 128 banks, global batch 8192): weak scaling, value = N x global steps/s = "config-2 units per second".
       Before anything is timed a parity preflight compares the N-rank sharded step with the single-GPU CRDLoss on
       the same global batch (and the reference golden `crd_d128`); a mismatch aborts the run with rc != 0.

Beside it, in the same run and under the same clock sampler:
  roofline        gather kernel: algorithmic bytes / CUDA-event time inside the timed steps, vs the measured HBM copy peak
  roofline_kron   the tcgen05 Kronecker kernels (BASELINE configs 3, 4): forward, weight gradient, factor gradients and
                  forward+backward; algorithmic flops / CUDA-event time vs the TF32 peak (= measured bf16 peak / 2)
  fused_step      fusion -> CRD, forward + backward + Adam, as ONE CUDA graph: config 4 (TrilinearFusion_A 33^3 -> 96,
                  batch 8192, K 16384, n 1M) and config 1 (BilinearFusion 32x32 -> 64, batch 64), with their CPU baselines
  knn_positives   full-bank KNN positives of the stage-2 criterion (CRD_criterion_v10): 1024 anchors against 1M x 128 rows,
                  TF32 tcgen05 pass + exact re-score, vs the TF32 peak; sklearn on the host cores beside it
  kmeans_centres  per-class k-means centres of the same criterion (pos_extra "centers", num_pos > 2): one Lloyd iteration over
                  a 1M x 128 bank vs the HBM copy peak, the whole fit, sklearn KMeans on the host for one class beside it
  strong_scaling  BASELINE config 5 as stated: ONE problem (16M x 128 banks, global batch 8192) on N GPUs
  e2e*            `e2e` = host int64 contrast_idx uploaded every step (the reference's loader contract);
                  `e2e_device_idx` = contrast_idx drawn on the GPU inside the graph (InstanceSampler), `e2e_int32_idx`
  cpu_baseline    the reference's own modules (unmodified copies staged by `__graft_entry__.build()` under baseline/_ref/,
                  kind "reference") or, if absent, the oracle port (kind "port"), on all host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import contextlib
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "distill steps/s (CRD fwd/bwd, config-2 units)"
C2 = dict(B=1024, D=128, K=16384, n=1_000_000, s_dim=128, t_dim=128)
ROWS_PER_GPU_SHARDED = 2_000_000
C5 = dict(B=8192, D=128, K=16384, n=16_000_000, s_dim=128, t_dim=128)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def crd_algorithmic_bytes(B, K, D):
    """row gathers (both banks) + indices + row update r/w + v reads / dv writes."""
    gather = 2 * B * (K + 1) * D * 4
    idx = B * (K + 1) * 8 + B * 8
    upd = 2 * (2 * B * D * 4)
    vio = 4 * B * D * 4
    return gather + idx + upd + vio, gather + idx + vio


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  The sampler process is started
    before the warm-up (nvidia-smi needs a moment to attach on an 8-GPU box); `mark()` at the start of the timed region
    drops everything sampled before it."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.skip = 0
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "10", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def _lines(self):
        with open(self.f.name) as g:
            return g.read().splitlines()

    def wait_ready(self, timeout_s=10.0):
        t0 = time.perf_counter()
        while self.p is not None and time.perf_counter() - t0 < timeout_s and not self._lines():
            time.sleep(0.05)

    def mark(self):
        if self.p is not None:
            self.skip = len(self._lines())

    def summary(self, lines):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}

    def window(self):
        """Clock record since the last mark() (the sampler keeps running)."""
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.03)
        lines = self._lines()
        lines = lines[self.skip:] if len(lines) > self.skip + 1 else lines[-2:]
        return self.summary(lines)

    def stop(self):
        if self.p is None:
            return
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        os.unlink(self.f.name)


class EventTimer:
    """CUDA events on the launching (current) stream around the gather kernel's launches.  While a CUDA graph is being
    captured the events are created `external=True`: they become event-record nodes of the graph and every replay
    re-records them, so after the timed region each pair holds the kernel time of that graph's LAST replay."""

    def __init__(self):
        self.pairs = []
        self.enabled = False
        self._open = None

    def _event(self):
        return torch.cuda.Event(enable_timing=True, external=torch.cuda.is_current_stream_capturing())

    def start(self, name, dev):
        if self.enabled:
            e = self._event()
            e.record(torch.cuda.current_stream(dev))
            self._open = e

    def stop(self, name, dev):
        if self.enabled and self._open is not None:
            e = self._event()
            e.record(torch.cuda.current_stream(dev))
            self.pairs.append((self._open, e))
            self._open = None

    def mean_ms(self):
        return statistics.mean(a.elapsed_time(b) for a, b in self.pairs) if self.pairs else None


def workload_name(world):
    if world == 1:
        return "config 2: CRD step feat_dim=128 batch=1024 n_data=1M nce_k=16384, fwd+bwd+bank update"
    n, Bg = ROWS_PER_GPU_SHARDED * world, C2["B"] * world
    return (f"config-2 unit per GPU on the row-sharded bank: global batch {Bg}, n_data {n} "
            f"({ROWS_PER_GPU_SHARDED} rows/GPU), nce_k 16384" + (" = BASELINE config 5" if world == 8 else ""))


def make_config(world):
    """Identical for both arms (`--impl ours` / `--impl reference`)."""
    total_bytes, _ = crd_algorithmic_bytes(C2["B"], C2["K"], C2["D"])
    return {"workload": workload_name(world),
            "l2": "inputs exceed L2: 1.02 GB of bank rows gathered at random per GPU, 4 rotating 134 MB index sets; no explicit flush",
            "algorithmic_bytes_per_step": total_bytes}


def make_opt(cfg, n, **kw):
    return types.SimpleNamespace(s_dim=cfg["s_dim"], t_dim=cfg["t_dim"], feat_dim=cfg["D"], n_data=n,
                                 nce_k=cfg["K"], nce_t=0.07, nce_m=0.5, **kw)


def gen_inputs(cfg, B, n, gen, device, pinned=False):
    """One step's synthetic inputs: N(0,1) features, unique anchors, uniform negatives."""
    kw = dict(device=device, generator=gen)
    f_s = torch.randn(B, cfg["s_dim"], **kw)
    f_t = torch.randn(B, cfg["t_dim"], **kw)
    idx = torch.randperm(n, **kw)[:B].contiguous()
    cidx = torch.randint(0, n, (B, cfg["K"] + 1), **kw)
    cidx[:, 0] = idx
    out = (f_s, f_t, idx, cidx)
    if pinned:
        out = tuple(t.pin_memory() for t in out)
    return out


# ------------------------------------------------------------------------------------------ #
# CPU arm: the reference's own modules (baseline/_ref) or the oracle port, on the host cores
# ------------------------------------------------------------------------------------------ #
_REF = None


def load_reference():
    """The UNMODIFIED reference sources staged under baseline/_ref/ by `__graft_entry__.build()` (git-ignored copies of
    MICCAI-2022/fusion.py, CL_utils/CRD_criterion.py), imported with the three shims of SURVEY.md §8(c):
      1. a stub `utils` module exposing init_max_weights with the semantics of utils.py:239-244 (the real utils.py imports
         lifelines / imblearn / torch_geometric, none installed);
      2. `torch.cuda.FloatTensor -> torch.FloatTensor` while a fusion forward runs on the host (fusion.py:56-57);
      3. `AliasMethod.cuda -> no-op` (CRD_criterion.py:17 calls it in the constructor).
    Plus, for n_data >= 65536 only: the O(n) python loop of AliasMethod.__init__ (33 us per row, 35 s at n = 1M, never on the
    timed path) is replaced by its closed form for uniform weights (prob = 1, alias = 0; bit-exact, tests/test_oracle_golden).
    Returns None when the copies are absent (the GPU box of a run that did not go through build())."""
    global _REF
    if _REF is not None:
        return _REF or None
    paths = {"fusion": os.path.join(REF_DIR, "fusion.py"), "crd": os.path.join(REF_DIR, "CL_utils", "CRD_criterion.py")}
    if not all(os.path.exists(p) for p in paths.values()):
        _REF = False
        return None
    import math
    stub = types.ModuleType("utils")

    def init_max_weights(module):
        for m in module.modules():
            if type(m) == torch.nn.Linear:
                stdv = 1. / math.sqrt(m.weight.size(1))
                m.weight.data.normal_(0, stdv)
                m.bias.data.zero_()
    stub.init_max_weights = init_max_weights
    saved = sys.modules.get("utils")
    sys.modules["utils"] = stub
    mods = {}
    try:
        for name, path in paths.items():
            spec = importlib.util.spec_from_file_location(f"_mml_ref_{name}", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mods[name] = mod
    finally:
        if saved is not None:
            sys.modules["utils"] = saved
        else:
            sys.modules.pop("utils", None)
    crd = mods["crd"]
    crd.AliasMethod.cuda = lambda self: None
    orig_init = crd.AliasMethod.__init__

    def alias_init(self, probs):
        if len(probs) >= 65536 and bool((probs == probs[0]).all()):
            if probs.sum() > 1:
                probs.div_(probs.sum())
            self.prob = torch.ones(len(probs))
            self.alias = torch.zeros(len(probs), dtype=torch.long)
        else:
            orig_init(self, probs)
    crd.AliasMethod.__init__ = alias_init
    _REF = types.SimpleNamespace(**mods)
    return _REF


@contextlib.contextmanager
def host_float_tensor():
    saved = torch.cuda.FloatTensor
    torch.cuda.FloatTensor = torch.FloatTensor
    try:
        yield
    finally:
        torch.cuda.FloatTensor = saved


def cpu_crd_steps_per_s(cfg, steps, warmup, sample_B, seed=2019):
    """Times `steps` CRD steps (forward, backward, Adam on the heads) of the reference's CRDLoss -- or of the oracle port
    when baseline/_ref is absent -- on a slice of `sample_B` anchors at full K / n_data and scales to the full batch (the
    work is linear in anchors).  Returns (full-workload steps/s, seconds per sample step, kind)."""
    torch.manual_seed(seed)
    n, D, K = cfg["n"], cfg["D"], cfg["K"]
    ref = load_reference()
    gen = torch.Generator().manual_seed(seed)
    times = []
    if ref is not None:
        crit = ref.crd.CRDLoss(make_opt(cfg, n))
        optim = torch.optim.Adam(list(crit.embed_s.parameters()) + list(crit.embed_t.parameters()), lr=2e-4, betas=(0.9, 0.999))

        def step(f_s, f_t, idx, cidx):
            loss = crit(f_s, f_t, idx, cidx)
            loss.backward()
            optim.step()
            optim.zero_grad(set_to_none=True)
        kind = "reference"
    else:
        from oracle import crd_oracle as co
        sd = {}
        for side, dim in (("s", cfg["s_dim"]), ("t", cfg["t_dim"])):
            l0, l2 = torch.nn.Linear(dim, D), torch.nn.Linear(D, D)
            sd[f"embed_{side}.linear.0.weight"], sd[f"embed_{side}.linear.0.bias"] = l0.weight.detach(), l0.bias.detach()
            sd[f"embed_{side}.linear.2.weight"], sd[f"embed_{side}.linear.2.bias"] = l2.weight.detach(), l2.bias.detach()
        for k in list(sd):
            sd[k].requires_grad_(True)
        optim = torch.optim.Adam([sd[k] for k in sd], lr=2e-4, betas=(0.9, 0.999))
        sd["contrast.memory_v1"], sd["contrast.memory_v2"] = co.memory_init(n, D)
        sd["contrast.params"] = co.make_params(K)

        def step(f_s, f_t, idx, cidx):
            loss, _, _ = co.crd_loss(sd, f_s, f_t, idx, cidx, n)
            loss.backward()
            optim.step()
            optim.zero_grad(set_to_none=True)
        kind = "port"
    for i in range(warmup + steps):
        f_s, f_t, idx, cidx = gen_inputs(cfg, sample_B, n, gen, "cpu")
        f_s.requires_grad_(True)
        t0 = time.perf_counter()
        step(f_s, f_t, idx, cidx)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t_sample = statistics.mean(times)
    return 1.0 / (t_sample * cfg["B"] / sample_B), t_sample, kind


def cpu_fused_step(which, steps, warmup, sample_B=None):
    """CPU baseline of a fused fusion -> CRD step: config 1 in full; config 4 with the fusion at the full batch and the
    CRD part on a `sample_B`-anchor slice scaled to 8192 anchors.  -> dict."""
    ref = load_reference()
    torch.manual_seed(2019)
    if which == "c1":
        B, dims, N, D, K, n = 64, (32, 32), 64, 128, 4096, 4096
    else:
        B, dims, N, D, K, n = 8192, (32, 32, 32), 96, 128, 16384, 1_000_000
    cfg = dict(B=B, D=D, K=K, n=n, s_dim=N, t_dim=N)
    Bc = B if sample_B is None else sample_B
    if ref is not None:
        with host_float_tensor():
            fusion = (ref.fusion.BilinearFusion(skip=0, dim1=32, dim2=32, mmhid=64, dropout_rate=0.25) if which == "c1" else
                      ref.fusion.TrilinearFusion_A(skip=1, dim1=32, dim2=32, dim3=32, mmhid=96)).train()
        crit = ref.crd.CRDLoss(make_opt(cfg, n))
        params = list(fusion.parameters()) + list(crit.embed_s.parameters()) + list(crit.embed_t.parameters())
        kind = "reference"
    else:
        import multimodal_learning_b200 as pkg
        from oracle import crd_oracle as co
        from oracle import fusion_oracle as fo
        mod = (pkg.BilinearFusion(skip=0, dim1=32, dim2=32, mmhid=64, dropout_rate=0.0) if which == "c1" else
               pkg.TrilinearFusion_A(skip=1, dim1=32, dim2=32, dim3=32, mmhid=96))
        sd_f = {k: v.clone() for k, v in mod.state_dict().items()}
        sd_c = {k: v.clone() for k, v in pkg.CRDLoss(make_opt(cfg, n)).state_dict().items()}
        params = []
        for sd in (sd_f, sd_c):
            for k in sd:
                if sd[k].is_floating_point() and "running" not in k and "memory" not in k and "params" not in k:
                    sd[k].requires_grad_(True)
                    params.append(sd[k])
        kind = "port"
    optim = torch.optim.Adam(params, lr=2e-4)
    gen = torch.Generator().manual_seed(5)
    t_fus, t_crd = [], []
    for i in range(warmup + steps):
        vecs = [torch.randn(B, d, generator=gen) for d in dims]
        f_t = torch.randn(Bc, N, generator=gen)
        idx = torch.randperm(n, generator=gen)[:Bc]
        cidx = torch.randint(0, n, (Bc, K + 1), generator=gen)
        cidx[:, 0] = idx
        t0 = time.perf_counter()
        if kind == "reference":
            with host_float_tensor():
                f = fusion(*vecs)
        elif which == "c1":
            f = fo.bilinear_fusion_forward(sd_f, *vecs, skip=0, training=True)
        else:
            f = fo.trilinear_fusion_forward(sd_f, *vecs, variant="A", skip=1)
        t1 = time.perf_counter()
        if kind == "reference":
            loss = crit(f[:Bc], f_t, idx, cidx)
        else:
            loss, _, _ = co.crd_loss(sd_c, f[:Bc], f_t, idx, cidx, n)
        if Bc != B:
            loss = loss + 0.0 * f[Bc:].sum()          # the whole batch goes through the fusion backward
        t2 = time.perf_counter()
        with (host_float_tensor() if kind == "reference" else contextlib.nullcontext()):
            loss.backward()
        optim.step()
        optim.zero_grad(set_to_none=True)
        t3 = time.perf_counter()
        if i >= warmup:
            # backward time is split in proportion to the forward times of the two parts
            fw_f, fw_c = t1 - t0, t2 - t1
            bw = t3 - t2
            t_fus.append(fw_f + bw * fw_f / (fw_f + fw_c))
            t_crd.append(fw_c + bw * fw_c / (fw_f + fw_c))
    tf, tc = statistics.mean(t_fus), statistics.mean(t_crd)
    total = tf + tc * (B / Bc)
    sample = (f"{steps} full steps" if Bc == B else
              f"{steps} steps: fusion at the full batch {B} ({tf * 1e3:.0f} ms), CRD on {Bc} of {B} anchors at full K/n "
              f"({tc * 1e3:.0f} ms) scaled x{B // Bc}")
    return {"value": 1.0 / total, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(host_cores())          # torchrun exports OMP_NUM_THREADS=1: undo it, this arm owns the host
    cfg = dict(C2)
    sample_B = 64
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sys.stderr):
        sps, t_sample, kind = cpu_crd_steps_per_s(cfg, args.steps, args.warmup, sample_B)
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / sps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.gpus),
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{sample_B} of 1024 anchors per step at full K=16384/n=1M "
                                   f"({t_sample * 1e3:.0f} ms each), scaled x{cfg['B'] // sample_B} (work is linear in anchors)",
                         "extrapolation_factor": cfg["B"] // sample_B,
                         "note": "one config-2 unit (1024 anchors x 16385 columns, n_data 1M) per step on the host cores; at "
                                 "N > 1 the GPU arm runs N such units per step, the CPU arm one: both report units/s"},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ #
# GPU arm
# ------------------------------------------------------------------------------------------ #
def _time_ms(fn, iters, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def knn_section(dev, peaks, cpu_anchors):
    """N4 of SURVEY 8(f): the full-bank KNN positives of `MIA 2023/.../CRD_criterion_v10.py:69-80` at config 2's bank
    (1M x 128, 1024 anchors, 5 positives, 3 classes): whole call (inverse norms, query prep, TF32 tcgen05 pass with the
    candidate filter, exact re-score, exact scan of flagged anchors) timed with CUDA events; beside it the reference's
    own way (sklearn cosine_similarity + sort on the host cores) on a slice of anchors."""
    from multimodal_learning_b200 import crd_knn
    n, D, B, P = 1_000_000, 128, 1024, 5
    gen = torch.Generator(device=dev).manual_seed(0)
    bank = torch.randn(n, D, device=dev, generator=gen)
    bank = bank / bank.norm(dim=1, keepdim=True)
    labels = torch.randint(0, 3, (n,), device=dev, generator=gen, dtype=torch.int32)
    rows = torch.randperm(n, device=dev, generator=gen)[:B]
    blab = labels[rows].long()
    for _ in range(3):
        idx, sim, flags = crd_knn.knn_positives(bank, labels, rows, blab, P, n_classes=3, return_flags=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        crd_knn.knn_positives(bank, labels, rows, blab, P, n_classes=3)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    inv = crd_knn.knn_inv_norms(bank)                    # what the module keeps between steps (only the batch's rows change)
    e0.record()
    for _ in range(iters):
        crd_knn.knn_positives(bank, labels, rows, blab, P, n_classes=3, inv_norms=inv)
    e1.record()
    torch.cuda.synchronize()
    ms_cached = e0.elapsed_time(e1) / iters
    tf32_peak = float(peaks["bf16_tflops"]) / 2
    flops = 2.0 * B * n * D
    out = {"workload": f"KNN positives over the full bank: {n} x {D} rows, {B} anchors, num_pos {P}, 3 classes",
           "ms": ms, "ms_with_cached_norms": ms_cached, "bound": "tensor", "achieved": flops / ms / 1e9, "peak": tf32_peak, "unit": "TFLOP/s",
           "frac": flops / ms / 1e9 / tf32_peak, "flops": flops, "own_kernel_launches": 9,
           "bank_passes_GBps": 2 * n * D * 4 / ms / 1e6, "flagged_anchors": int(flags.sum()),
           "note": "time of the whole call (9 launches: norms, queries, sampling pass, floor, full pass, re-score, flagged-anchor scan x3); flops = 2 B n D of the TF32 pass; the bank is read once for the "
                   "inverse norms and once by TMA for the GEMM (L2-shared between the 8 anchor tiles)"}
    if cpu_anchors > 0:
        try:
            from sklearn.metrics.pairwise import cosine_similarity
            torch.set_num_threads(host_cores())
            hb, hl = bank.cpu(), labels.cpu().long()
            hr, hlab = rows[:cpu_anchors].cpu(), blab[:cpu_anchors].cpu()
            t0 = time.perf_counter()
            mask = (hl.view(1, -1) == hlab.view(-1, 1)).float()
            sc = mask * torch.tensor(cosine_similarity(hb[hr].numpy(), hb.numpy()))
            order = torch.sort(sc, descending=True, dim=-1)[1][:, :P]
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": dt * 1e3 * B / cpu_anchors, "unit": "ms per call", "cores": torch.get_num_threads(),
                                   "kind": "reference", "same_neighbours": bool(torch.equal(order.to(dev), idx[:cpu_anchors])),
                                   "sample": f"{cpu_anchors} of {B} anchors: sklearn cosine_similarity against all {n} rows, class "
                                             f"mask, torch.sort (CRD_criterion_v10.py:69-74), scaled x{B // cpu_anchors}"}
        except Exception as e:                       # noqa: BLE001
            out["cpu_baseline_error"] = f"{type(e).__name__}: {e}"
    return out


def kmeans_section(dev, peaks, cpu):
    """N4, pos_extra == "centers" with num_pos > 2 (`MIA 2023/.../CRD_criterion_v10.py:84-92`): per-class k-means centres of
    a 1M x 128 bank, 3 classes, 3 centres each.  One Lloyd iteration (assign pass over the bank + update) timed with CUDA
    events against the measured HBM copy peak, the whole fit (tolerance, k-means++ start, iterations until every class
    stops) by wall clock; beside it the reference's own way for ONE class (rows copied to the host, sklearn KMeans)."""
    from multimodal_learning_b200 import crd_kmeans as km
    n, D, k = 1 << 20, 128, 3
    gen = torch.Generator(device=dev).manual_seed(0)
    modes = torch.randn(24, D, device=dev, generator=gen)
    bank = torch.nn.functional.normalize(modes[torch.randint(0, 24, (n,), device=dev, generator=gen)]
                                         + 0.5 * torch.randn(n, D, device=dev, generator=gen), dim=1).contiguous()
    labels = torch.randint(0, 3, (n,), device=dev, generator=gen)
    class_idx = [torch.nonzero(labels == c).flatten().cpu().numpy() for c in range(3)]
    cls = km.ClassRows(class_idx, dev)
    gen.manual_seed(1)
    centres = km.kmeans_plus_plus(bank, cls, k, gen)
    ws = km.lloyd(bank, cls, centres, iterations=3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    km.lloyd(bank, cls, centres, iterations=reps, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    it_ms = e0.elapsed_time(e1) / reps
    gen.manual_seed(1)
    km.class_kmeans(bank, cls, k, generator=gen)
    gen.manual_seed(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit, info = km.class_kmeans(bank, cls, k, generator=gen, return_info=True)
    torch.cuda.synchronize()
    fit_ms = (time.perf_counter() - t0) * 1e3
    nbytes = n * (4 * D + 8)
    peak = float(peaks["hbm_gbs"])
    out = {"workload": f"per-class k-means centres: {n} x {D} bank, 3 classes, {k} centres per class",
           "lloyd_iteration_ms": it_ms, "fit_ms": fit_ms, "iterations_enqueued": info["iterations_enqueued"],
           "converged": bool(info["done"].all()), "bound": "hbm", "achieved": nbytes / it_ms / 1e6, "peak": peak, "unit": "GB/s",
           "frac": nbytes / it_ms / 1e6 / peak, "bytes": nbytes, "own_kernel_launches": 2,
           "note": "one iteration = kmeans_assign_kernel (every listed row read once: 4 D + 8 bytes) + kmeans_update_kernel; "
                   "the assign pass is bound by shared-memory wavefronts, see profiles/r2_kmeans_ncu.txt"}
    if cpu:
        try:
            from sklearn.cluster import KMeans
            t0 = time.perf_counter()
            X = bank.index_select(0, cls.rows[:cls.sizes[0]]).cpu().numpy()
            est = KMeans(n_clusters=k).fit(X)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": dt * 1e3 * 3, "unit": "ms per bank", "cores": host_cores(), "kind": "reference",
                                   "iterations": int(est.n_iter_),
                                   "sample": f"class 0 only ({X.shape[0]} rows): index_select -> .cpu().numpy() -> sklearn KMeans(n_clusters={k}).fit "
                                             "(CRD_criterion_v10.py:86-92), scaled x3 classes"}
        except Exception as e:                       # noqa: BLE001
            out["cpu_baseline_error"] = f"{type(e).__name__}: {e}"
    return out


def kron_section(dev, peaks, peak_kind):
    """BASELINE configs 3 and 4: the tcgen05 Kronecker kernels through the public functional `kron_linear` (forward;
    weight gradient = transposes + K3 + unpack; factor gradients = transpose + K2 + reduce), each timed alone with CUDA
    events over 20 launches after 5 warm-up launches.  One call streams 0.3-1.1 GB of weights / activations plus a fresh
    workspace, more than the 126 MB L2.  Peak: measured bf16 burst / 2 (TF32 operands)."""
    from multimodal_learning_b200.fusion import KronLinearState, kron_linear
    tf32_peak = float(peaks["bf16_tflops"]) / 2
    points = [("c3 BilinearFusion d=128 N=256 B=65536", 65536, (128, 128), 256, 0.0),
              ("c3 BilinearFusion d=128 N=256 B=16384", 16384, (128, 128), 256, 0.0),
              ("c3 BilinearFusion d=64 N=128 B=65536", 65536, (64, 64), 128, 0.0),
              ("c4 TrilinearFusion 33^3 N=96 B=8192", 8192, (32, 32, 32), 96, 0.0),
              ("c4 TrilinearFusion 33^3 N=96 B=8192, post_fusion_dropout 0.25 (module default)", 8192, (32, 32, 32), 96, 0.25)]
    out = []
    for name, B, dims, N, p in points:
        kk = 1
        for d in dims:
            kk *= d + 1
        gen = torch.Generator(device=dev).manual_seed(0)
        fs = [torch.rand(B, d, device=dev, generator=gen) for d in dims]
        W = torch.randn(N, kk, device=dev, generator=gen) / kk ** 0.5
        bias = torch.zeros(N, device=dev)
        st = KronLinearState(dims)
        tr = p > 0
        flops = 2.0 * B * kk * N
        fsg = [f.clone().requires_grad_(True) for f in fs]
        Wg = W.clone().requires_grad_(True)
        y_all = kron_linear(st, fsg, Wg, bias, drop_p=p, training=tr, seed=7)
        y_w = kron_linear(st, fs, Wg, bias, drop_p=p, training=tr, seed=7)
        y_f = kron_linear(st, fsg, W, bias, drop_p=p, training=tr, seed=7)
        G = torch.randn_like(y_all)
        ms = {
            "fwd": _time_ms(lambda: kron_linear(st, fs, W, bias, drop_p=p, training=tr, seed=7), 20, 5),
            "wgrad": _time_ms(lambda: torch.autograd.grad(y_w, [Wg], G, retain_graph=True), 10, 3),
            "dgrad": _time_ms(lambda: torch.autograd.grad(y_f, fsg, G, retain_graph=True), 10, 3),
            "bwd": _time_ms(lambda: torch.autograd.grad(y_all, [Wg] + fsg, G, retain_graph=True), 10, 3),
        }
        rec = {"point": name, "B": B, "dims": list(dims), "N": N, "Kk": kk, "dropout": p, "bound": "tensor",
               "flops_fwd": flops, "unit": "TFLOP/s", "peak": tf32_peak,
               "peak_source": f"MEASURED_PEAKS.json ({peak_kind}) bf16_tflops / 2 (TF32 operands), burst"}
        for k, mult in (("fwd", 1.0), ("wgrad", 1.0), ("dgrad", 1.0), ("bwd", 2.0)):
            a = mult * flops / ms[k] / 1e9
            rec[k] = {"ms": ms[k], "achieved": a, "frac": a / tf32_peak}
        fb = 3.0 * flops / (ms["fwd"] + ms["bwd"]) / 1e9
        rec["fwd_bwd"] = {"ms": ms["fwd"] + ms["bwd"], "achieved": fb, "frac": fb / tf32_peak}
        out.append(rec)
        del fs, fsg, W, Wg, y_all, y_w, y_f, G
        torch.cuda.empty_cache()
    return out


def fused_step_section(dev, which, steps, warmup):
    """fusion -> CRD, forward + backward + Adam over fusion + Embed parameters, whole step replayed as ONE
    CUDA graph per rotating input set; inputs resident on the device."""
    import multimodal_learning_b200 as pkg
    torch.manual_seed(2019)
    if which == "c1":
        name = "config 1: BilinearFusion(32,32->64, dropout 0.25) -> CRDLoss, batch 64, n_data 4096, nce_k 4096"
        B, dims, N, D, K, n = 64, (32, 32), 64, 128, 4096, 4096
        fusion = pkg.BilinearFusion(skip=0, dim1=32, dim2=32, mmhid=64, dropout_rate=0.25)
    else:
        name = ("config 4: TrilinearFusion_A(32,32,32->96, post-fusion dropout 0.25; 33^3 outer product never materialised) "
                "-> CRDLoss, batch 8192, n_data 1M, nce_k 16384")
        B, dims, N, D, K, n = 8192, (32, 32, 32), 96, 128, 16384, 1_000_000
        fusion = pkg.TrilinearFusion_A(skip=1, dim1=32, dim2=32, dim3=32, mmhid=96)
    fusion = fusion.to(dev).train()
    cfg = dict(B=B, D=D, K=K, n=n, s_dim=N, t_dim=N)
    crd = pkg.CRDLoss(make_opt(cfg, n)).to(dev)
    params = list(fusion.parameters()) + list(crd.parameters())
    optim = torch.optim.Adam(params, lr=2e-4, betas=(0.9, 0.999), fused=True)
    gen = torch.Generator(device=dev).manual_seed(7)
    pool = []
    for _ in range(2 if which == "c4" else 4):
        vecs = [torch.randn(B, d, device=dev, generator=gen) for d in dims]
        f_t = torch.randn(B, N, device=dev, generator=gen)
        idx = torch.randperm(n, device=dev, generator=gen)[:B].contiguous()
        cidx = torch.randint(0, n, (B, K + 1), device=dev, generator=gen)
        cidx[:, 0] = idx
        pool.append((vecs, f_t, idx, cidx))
    for i in range(3):
        vecs, f_t, idx, cidx = pool[i % len(pool)]
        for p in params:
            p.grad = None
        loss = crd(fusion(*vecs), f_t, idx, cidx)
        loss.backward()
        optim.step()
    del loss
    l0 = pkg._cabi.launch_count()
    optim = torch.optim.Adam(params, lr=2e-4, betas=(0.9, 0.999), fused=True, capturable=True)
    nv = len(dims)
    gstep = pkg.GraphedTrainStep(lambda *a: crd(fusion(*a[:nv]), *a[nv:]), params, optim, (*pool[0][0], *pool[0][1:]),
                                 warmup=2, n_buffers=len(pool))
    launches_per_step = (pkg._cabi.launch_count() - l0) // (len(pool) + 2)
    for slot, (vecs, f_t, idx, cidx) in enumerate(pool):
        for dst, src in zip(gstep.buffers(slot), (*vecs, f_t, idx, cidx)):
            dst.detach().copy_(src)
    ms = _time_ms(lambda: gstep.replay(), steps, max(warmup, 3))
    kk = 1
    for d in dims:
        kk *= d + 1
    total_bytes, _ = crd_algorithmic_bytes(B, K, D)
    rec = {"workload": name, "value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms, "steps": steps,
           "launch_mode": "cuda_graph", "own_kernel_launches_per_step": launches_per_step,
           "kron_flops_fwd_bwd": 6.0 * B * kk * N, "crd_algorithmic_bytes": total_bytes,
           "loss_finite": bool(torch.isfinite(gstep.static_loss[0]).all().item())}
    del gstep, pool, crd, fusion, optim, params
    torch.cuda.empty_cache()
    return rec


def preflight_parity(world, rank, dev, dist):
    """N > 1: the sharded step must reproduce the single-GPU step.  (1) reference golden `crd_d128` through ShardedCRDLoss;
    (2) a bank of 262144 rows, K = 4096, global batch 64 x world: loss, Embed-parameter gradients, local dL/df_s and the
    updated banks of the N-rank step vs the single-GPU CRDLoss on the same global batch (run on every rank).  Raises."""
    import multimodal_learning_b200 as pkg
    from multimodal_learning_b200.sharded import ShardedCRDLoss

    def rel(a, b):
        d = b.abs().max().item()
        return (a - b).abs().max().item() / (d if d > 0 else 1.0)
    worst = 0.0
    cfg = dict(B=64 * world, D=128, K=4096, n=262144, s_dim=128, t_dim=128)
    torch.manual_seed(99)
    single = pkg.CRDLoss(make_opt(cfg, cfg["n"], bank_device=dev)).to(dev)
    for t in list(single.parameters()) + [single.contrast.memory_v1, single.contrast.memory_v2]:
        dist.broadcast(t.data, src=0)
    shard = ShardedCRDLoss(make_opt(cfg, cfg["n"]), device=dev)
    shard.embed_s.load_state_dict(single.embed_s.state_dict())
    shard.embed_t.load_state_dict(single.embed_t.state_dict())
    shard.contrast.load_full_banks(single.contrast.memory_v1, single.contrast.memory_v2)
    gen = torch.Generator(device=dev).manual_seed(4242)          # same stream on every rank: the GLOBAL batch
    Bl = cfg["B"] // world
    sl = slice(rank * Bl, (rank + 1) * Bl)
    for step in range(2):
        f_s, f_t, idx, cidx = gen_inputs(cfg, cfg["B"], cfg["n"], gen, dev)
        a = f_s.clone().requires_grad_(True)
        single.zero_grad()
        with contextlib.redirect_stdout(sys.stderr):
            want = single(a, f_t, idx, cidx)
        want.backward()
        b = f_s[sl].clone().requires_grad_(True)
        shard.zero_grad()
        with contextlib.redirect_stdout(sys.stderr):
            got = shard(b, f_t[sl].contiguous(), idx[sl].contiguous(), cidx[sl].contiguous())
        got.backward()
        errs = {"loss": rel(got, want), "grad_f_s": rel(b.grad, a.grad[sl])}
        for (k, p), (_, q) in zip(shard.named_parameters(), single.named_parameters()):
            errs["grad." + k] = rel(p.grad, q.grad)
        m1, m2 = shard.contrast.gather_full_banks()
        errs["memory_v1"] = rel(m1, single.contrast.memory_v1)
        errs["memory_v2"] = rel(m2, single.contrast.memory_v2)
        for k, e in errs.items():
            if not e < 1e-4:
                raise RuntimeError(f"sharded-vs-single parity FAILED on rank {rank}, step {step}, {k}: rel {e:.3e} (tolerance 1e-4)")
            worst = max(worst, e)
    del single, shard
    torch.cuda.empty_cache()
    # (1) the reference-generated golden steps (tests/golden/crd_d128.npz: B=16, D=128, K=128, n=300, 2 steps)
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "crd_d128.npz"))
    gcfg = json.loads(bytes(z["__config__"]).decode())

    def gt(key):
        return torch.from_numpy(z[key].copy()).to(dev)
    if gcfg["B"] % world == 0 and (gcfg["B"] // world) % 2 == 0:
        gopt = types.SimpleNamespace(s_dim=gcfg["s_dim"], t_dim=gcfg["t_dim"], feat_dim=gcfg["D"], n_data=gcfg["n"], nce_k=gcfg["K"],
                                     nce_t=0.07, nce_m=0.5)
        gm = ShardedCRDLoss(gopt, device=dev)
        gm.embed_s.load_state_dict({k[len("init.embed_s."):]: gt(k) for k in z.files if k.startswith("init.embed_s.")})
        gm.embed_t.load_state_dict({k[len("init.embed_t."):]: gt(k) for k in z.files if k.startswith("init.embed_t.")})
        gm.contrast.load_full_banks(gt("init.contrast.memory_v1"), gt("init.contrast.memory_v2"))
        gl = gcfg["B"] // world
        gs = slice(rank * gl, (rank + 1) * gl)
        for step in range(gcfg["steps"]):
            pfx = f"step{step}."
            a = gt(pfx + "f_s")[gs].clone().requires_grad_(True)
            gm.zero_grad()
            with contextlib.redirect_stdout(sys.stderr):
                got = gm(a, gt(pfx + "f_t")[gs].contiguous(), gt(pfx + "idx")[gs].contiguous(), gt(pfx + "contrast_idx")[gs].contiguous())
            got.backward()
            m1, _ = gm.contrast.gather_full_banks()
            for k, e in (("loss", rel(got, gt(pfx + "loss"))), ("grad_f_s", rel(a.grad, gt(pfx + "grad_f_s")[gs])),
                         ("memory_v1", rel(m1, gt(pfx + "memory_v1")))):
                if not e < 1e-4:
                    raise RuntimeError(f"golden crd_d128 parity FAILED on rank {rank}, step {step}, {k}: rel {e:.3e}")
                worst = max(worst, e)
        del gm
        torch.cuda.empty_cache()
    return {"checked": True, "world": world, "bank_rows": cfg["n"], "nce_k": cfg["K"], "global_batch": cfg["B"], "steps": 2,
            "tolerance": 1e-4, "worst_rel": worst,
            "what": "N-rank ShardedCRDLoss vs single-GPU CRDLoss on the same global batch (loss, dL/df_s, Embed gradients, updated "
                    "banks) and vs the reference-generated golden steps tests/golden/crd_d128.npz"}


def run_gpu_arm(args):
    import torch.distributed as dist

    import multimodal_learning_b200 as pkg
    from multimodal_learning_b200 import crd as crd_mod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(C2)
    B, D, K = cfg["B"], cfg["D"], cfg["K"]
    pkg._cabi.lib()
    peaks, peak_kind = measured_peaks()
    errors = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    parity = None
    if world > 1:
        parity = preflight_parity(world, rank, dev, dist)      # raises (rc != 0) on mismatch
        barrier()

    torch.manual_seed(2019)
    if world == 1:
        n = cfg["n"]
        mod = pkg.CRDLoss(make_opt(cfg, n)).to(dev)
        B_local = B
    else:
        from multimodal_learning_b200.sharded import ShardedCRDLoss
        n = ROWS_PER_GPU_SHARDED * world
        B_local = B
        mod = ShardedCRDLoss(make_opt(cfg, n), device=dev)
    opt_params = [p for p in mod.parameters()]
    optim = torch.optim.Adam(opt_params, lr=2e-4, betas=(0.9, 0.999), fused=True)   # networks_new.py:85

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    pool = [gen_inputs(cfg, B_local, n, gen, dev) for _ in range(4)]
    hgen = torch.Generator().manual_seed(4321 + rank)
    host_pool = [gen_inputs(cfg, B_local, n, hgen, "cpu", pinned=True) for _ in range(3)]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_pool[0])

    def step_device(i):
        f_s, f_t, idx, cidx = pool[i % len(pool)]
        f_s = f_s.detach().requires_grad_(True)
        for p in opt_params:
            p.grad = None
        loss = mod(f_s, f_t, idx, cidx)
        loss.backward()
        optim.step()
        return loss

    profile_region = os.environ.get("MML_BENCH_PROFILE") == "1"   # ncu --profile-from-start off

    def timed(fn, steps, profile=False):
        barrier()
        if profile and profile_region:
            torch.cuda.profiler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        if profile and profile_region:
            torch.cuda.profiler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- device-resident throughput (`value`) ----
    # Default: the whole step (Embed heads, routing / gather / finishers / row update, backward, Adam; on N > 1 also the
    # symmetric-memory barriers and the peer-memory gradient reduction) is captured ONCE per input set into a CUDA graph
    # (`GraphedTrainStep`) and replayed: one launch per step instead of ~65.  `--eager` times the eager launches instead.
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for i in range(max(args.warmup, 3)):
        step_device(i)
    timer = EventTimer()
    mode = "eager"
    routing_mode = None
    captured_launches = None
    step_fn = step_device
    if not args.eager:
        try:
            barrier()
            g_optim = torch.optim.Adam(opt_params, lr=2e-4, betas=(0.9, 0.999), fused=True, capturable=True)
            mark = {}

            def before_capture():
                crd_mod.KERNEL_TIMER = timer
                timer.enabled = True
                mark["l0"] = pkg._cabi.launch_count()
            if world > 1 and len(pool) % 2 == 0 and os.environ.get("MML_PREFETCH_ROUTING", "1") == "1":
                # the next resident batch's indices are routed under this step's tail (ShardedCRDLoss next_contrast_idx)
                gstep = pkg.GraphedTrainStep(
                    lambda a, b, c, d, next_inputs=None: mod(a, b, c, d, next_contrast_idx=next_inputs[3]), opt_params, g_optim,
                    pool[0], grad_inputs=(0,), warmup=3, n_buffers=len(pool), before_capture=before_capture, pass_next_inputs=True)
                routing_mode = "prefetched: the next step's indices are routed under this step's tail"
            else:
                gstep = pkg.GraphedTrainStep(lambda a, b, c, d: mod(a, b, c, d), opt_params, g_optim, pool[0], grad_inputs=(0,),
                                             warmup=3, n_buffers=len(pool), before_capture=before_capture)
                routing_mode = "in step" if world > 1 else None
            timer.enabled = False
            crd_mod.KERNEL_TIMER = None
            captured_launches = (pkg._cabi.launch_count() - mark["l0"]) // len(pool)
            for slot, entry in enumerate(pool):
                for dst, src in zip(gstep.buffers(slot), entry):
                    dst.detach().copy_(src)
            barrier()
            mode = "cuda_graph"
            step_fn = lambda i: gstep.replay()       # round-robin over the captured input sets  # noqa: E731
            for i in range(max(args.warmup, 3)):
                step_fn(i)
        except Exception as e:                       # noqa: BLE001 -- report and fall back to eager launches
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            timer = EventTimer()
            mode, step_fn = "eager", step_device
    if mode == "eager":
        crd_mod.KERNEL_TIMER = timer
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.wait_ready()
    for i in range(3):                      # the GPU is busy again when the region starts (all ranks: collectives inside)
        step_fn(i)
    launches0 = pkg._cabi.launch_count()
    if sampler:
        sampler.mark()
    if mode == "eager":
        timer.enabled = True
    total_ms = timed(step_fn, args.steps, profile=True)
    timer.enabled = False
    clocks = sampler.window() if sampler else None
    launches = (pkg._cabi.launch_count() - launches0) if mode == "eager" else captured_launches * args.steps
    crd_mod.KERNEL_TIMER = None
    ms_per_step = total_ms / args.steps
    value = world * 1000.0 / ms_per_step

    # ---- end to end through the public API with pinned HOST inputs ----
    copy_stream = torch.cuda.Stream(dev)
    staged = {}

    def stage(i, hp):
        with torch.cuda.stream(copy_stream):
            staged[i] = tuple(t.to(dev, non_blocking=True) for t in hp[i % len(hp)])
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            staged[i] += (ev,)

    def make_step_e2e(hp):
        def step_e2e(i):
            if i not in staged:
                stage(i, hp)
            f_s, f_t, idx, cidx, ev = staged.pop(i)
            torch.cuda.current_stream(dev).wait_event(ev)
            stage(i + 1, hp)                               # upload of step i+1 overlaps step i's kernels
            f_s = f_s.requires_grad_(True)
            for p in opt_params:
                p.grad = None
            loss = mod(f_s, f_t, idx, cidx)
            loss.backward()
            optim.step()
            for t in (f_s, f_t, idx, cidx):
                t.record_stream(torch.cuda.current_stream(dev))
            return loss.item()                             # D2H read of the step's result, every step
        return step_e2e

    step_e2e = make_step_e2e(host_pool)
    for i in range(3):
        step_e2e(i)
    staged.clear()
    e2e_eager_ms = timed(step_e2e, args.steps) / args.steps
    staged.clear()
    e2e_ms, e2e_note = e2e_eager_ms, "pinned host inputs; upload of step i+1 overlaps step i on a copy stream; eager launches"

    # Same step through `GraphedTrainStep` (the package's public whole-step CUDA graph): with a host sync per step the
    # eager path is bound by the host issuing ~65 launches, the graph is one launch.  Two captured graphs over two
    # static input sets: the H2D of step i+1 lands in the other set while step i replays.
    def graph_e2e(hp, example):
        g_opt = torch.optim.Adam(opt_params, lr=2e-4, betas=(0.9, 0.999), fused=True, capturable=True)
        gs = pkg.GraphedTrainStep(lambda a, b, c, d: mod(a, b, c, d), opt_params, g_opt, example, grad_inputs=(0,),
                                  warmup=3, n_buffers=2)
        barrier()
        up_done = [torch.cuda.Event(), torch.cuda.Event()]
        run_done = [torch.cuda.Event(), torch.cuda.Event()]

        def upload(i):
            slot = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(run_done[slot])             # the replay that last read this set has finished
                for dst, src in zip(gs.buffers(slot), hp[i % len(hp)]):
                    dst.detach().copy_(src, non_blocking=True)
                up_done[slot].record(copy_stream)
        primed = set()

        def step(i):
            slot = i % 2
            if i not in primed:
                upload(i)
            primed.discard(i)
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(up_done[slot])
            upload(i + 1)
            primed.add(i + 1)
            loss = gs.replay(slot)
            run_done[slot].record(cur)
            return loss.item()
        for ev in run_done:
            ev.record(torch.cuda.current_stream(dev))
        for i in range(4):
            step(i)
        primed.clear()
        barrier()
        ms = timed(step, args.steps) / args.steps
        primed.clear()
        return ms

    if mode == "cuda_graph":
        barrier()
        e2e_ms = graph_e2e(host_pool, pool[0])
        e2e_note = ("pinned host inputs -> static device buffers (H2D of step i+1 overlaps the replay of step i), whole step "
                    "replayed as one CUDA graph (GraphedTrainStep), loss read back every step")
    e2e_value = world * 1000.0 / e2e_ms

    # N3: contrast_idx produced ON THE DEVICE by the class-conditional sampler (what the reference's loader workers do with
    # numpy, data_loaders_MT.py:222-249) inside the captured step: only f_s, f_t, idx cross PCIe.
    e2e_device_idx = None
    if mode == "cuda_graph":
        try:
            barrier()
            labels = torch.randint(0, 3, (n,), generator=torch.Generator().manual_seed(7))
            sampler_mod = pkg.InstanceSampler(labels, nce_k=K).cuda(dev)
            s_optim = torch.optim.Adam(opt_params, lr=2e-4, betas=(0.9, 0.999), fused=True, capturable=True)
            sstep = pkg.GraphedTrainStep(lambda a, b, c: mod(a, b, c, sampler_mod(c)), opt_params, s_optim, pool[0][:3],
                                         grad_inputs=(0,), warmup=3, n_buffers=2)
            barrier()
            small_host = [hp[:3] for hp in host_pool]

            def step_sampler(i):
                for dst, src in zip(sstep.buffers(), small_host[i % len(small_host)]):
                    dst.detach().copy_(src, non_blocking=True)
                return sstep.replay().item()
            for i in range(4):
                step_sampler(i)
            barrier()
            sm_ms = timed(step_sampler, args.steps) / args.steps
            e2e_device_idx = {"value": world * 1000.0 / sm_ms, "unit": "steps/s", "ms_per_step": sm_ms,
                              "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in small_host[0]),
                              "d2h_bytes_per_step": 4,
                              "note": "contrast_idx drawn on the GPU by InstanceSampler (3 classes, exact positive, K negatives "
                                      "from the other classes; what the reference's DataLoader workers do with numpy) inside the "
                                      "replayed graph; f_s, f_t, idx uploaded from pinned host memory and the loss read back every step"}
            del sstep
        except Exception as e:                       # noqa: BLE001
            errors["e2e_device_idx"] = f"{type(e).__name__}: {e}"

    # same call with the caller handing contrast_idx over as int32 (n_data < 2^31): half the PCIe bytes, same graph path.
    e2e_i32 = None
    if mode == "cuda_graph" and world == 1:          # the sharded module routes int64 ids (the reference's dtype) only
        try:
            host_pool32 = [(hp[0], hp[1], hp[2], hp[3].to(torch.int32).pin_memory()) for hp in host_pool]
            ex32 = (pool[0][0], pool[0][1], pool[0][2], pool[0][3].to(torch.int32))
            i32_ms = graph_e2e(host_pool32, ex32)
            e2e_i32 = {"value": world * 1000.0 / i32_ms, "unit": "steps/s", "ms_per_step": i32_ms,
                       "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host_pool32[0]), "d2h_bytes_per_step": 4,
                       "note": "contrast_idx supplied as int32 by the caller; same two-graph pipeline as `e2e`"}
            del host_pool32, ex32
        except Exception as e:                       # noqa: BLE001
            errors["e2e_int32_idx"] = f"{type(e).__name__}: {e}"

    k_ms = timer.mean_ms()
    del pool, host_pool, mod, optim
    if mode == "cuda_graph":
        del gstep
    staged.clear()
    torch.cuda.empty_cache()

    # ---- BASELINE config 5 as stated: one problem (16M-row banks, global batch 8192) on N GPUs: strong scaling ----
    strong = None
    try:
        barrier()
        c5 = dict(C5)
        Bl5 = c5["B"] // world
        if world == 1:
            mod5 = pkg.CRDLoss(make_opt(c5, c5["n"], bank_device=dev)).to(dev)
        else:
            from multimodal_learning_b200.sharded import ShardedCRDLoss
            mod5 = ShardedCRDLoss(make_opt(c5, c5["n"]), device=dev)
        p5 = [p for p in mod5.parameters()]
        o5 = torch.optim.Adam(p5, lr=2e-4, betas=(0.9, 0.999), fused=True, capturable=True)
        g5 = torch.Generator(device=dev).manual_seed(777 + rank)
        pool5 = [gen_inputs(c5, Bl5, c5["n"], g5, dev) for _ in range(2)]
        with contextlib.redirect_stdout(sys.stderr):
            if world > 1 and routing_mode is not None and routing_mode.startswith("prefetched"):
                gs5 = pkg.GraphedTrainStep(lambda a, b, c, d, next_inputs=None: mod5(a, b, c, d, next_contrast_idx=next_inputs[3]),
                                           p5, o5, pool5[0], grad_inputs=(0,), warmup=3, n_buffers=2, pass_next_inputs=True)
            else:
                gs5 = pkg.GraphedTrainStep(lambda a, b, c, d: mod5(a, b, c, d), p5, o5, pool5[0], grad_inputs=(0,), warmup=3,
                                           n_buffers=2)
        for slot, entry in enumerate(pool5):
            for dst, src in zip(gs5.buffers(slot), entry):
                dst.detach().copy_(src)
        n5 = max(5, min(args.steps, 20))
        for i in range(3):
            gs5.replay()
        ms5 = timed(lambda i: gs5.replay(), n5) / n5
        tb5, _ = crd_algorithmic_bytes(c5["B"], c5["K"], c5["D"])
        strong = {"workload": "BASELINE config 5: row-sharded banks 16M x 128 fp32 x 2, nce_k 16384, GLOBAL batch 8192 "
                              f"({Bl5} anchors per GPU), fwd+bwd+bank update+Adam, CUDA graph",
                  "scaling": "strong", "n_gpus": world, "value": 1000.0 / ms5, "unit": "global steps/s", "ms_per_step": ms5,
                  "steps": n5, "algorithmic_bytes_per_global_step": tb5,
                  "hbm_frac_per_gpu": (tb5 / world) / (ms5 * 1e-3) / 1e9 / float(peaks["hbm_gbs"])}
        del gs5, pool5, mod5, o5, p5
        torch.cuda.empty_cache()
    except Exception as e:                           # noqa: BLE001
        errors["strong_scaling"] = f"{type(e).__name__}: {e}"

    # ---- the tensor half of the metric (N = 1: the Kronecker GEMM is data-parallel, weights replicated) ----
    kron, kron_clocks, fused = None, None, {}
    if world == 1 and not args.no_kron:
        try:
            if sampler:
                sampler.mark()
            kron = kron_section(dev, peaks, peak_kind)
            kron_clocks = sampler.window() if sampler else None
        except Exception as e:                       # noqa: BLE001
            errors["roofline_kron"] = f"{type(e).__name__}: {e}"
        for which, st in (("c4", max(5, min(args.steps, 20))), ("c1", max(50, args.steps * 5))):
            try:
                with contextlib.redirect_stdout(sys.stderr):
                    fused[which] = fused_step_section(dev, which, st, args.warmup)
            except Exception as e:                   # noqa: BLE001
                errors["fused_step." + which] = f"{type(e).__name__}: {e}"
    knn = None
    if world == 1 and not args.no_kron:
        try:
            knn = knn_section(dev, peaks, 0 if args.no_cpu else 8)
        except Exception as e:                       # noqa: BLE001
            errors["knn_positives"] = f"{type(e).__name__}: {e}"
    kmeans = None
    if world == 1 and not args.no_kron:
        try:
            kmeans = kmeans_section(dev, peaks, not args.no_cpu)
        except Exception as e:                       # noqa: BLE001
            errors["kmeans_centres"] = f"{type(e).__name__}: {e}"
    if sampler:
        sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_bytes, gather_bytes = crd_algorithmic_bytes(B, K, D)
    achieved = gather_bytes / (k_ms * 1e-3) / 1e9 if k_ms else None
    peak = float(peaks["hbm_gbs"])
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "launch_mode": mode,
        "config": make_config(world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": None,
                     "traffic_note": "not measured in this run; ncu --set full of this kernel at this config: dram read+write "
                                     "16.63 GB per launch = 0.96 x algorithmic (profiles/r2_crd_gather_full.txt)",
                     "kernel": "crd_gather_kernel<4,2,fused> (+2 finisher launches)", "kernel_ms": k_ms,
                     "timing": ("CUDA events recorded as external event nodes around the gather launch inside the replayed graphs "
                                "(last replay of each captured graph)" if mode == "cuda_graph" else
                                "CUDA events on the launching stream around every gather launch of the timed steps"),
                     "bytes_per_launch": gather_bytes, "peak_source": f"MEASURED_PEAKS.json ({peak_kind}, burst copy)"},
        "e2e": {"value": e2e_value, "unit": "steps/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "note": e2e_note, "eager_ms_per_step": e2e_eager_ms},
        "gpu_launches": launches, "clocks": clocks,
    }
    if routing_mode is not None:
        line["config"]["routing"] = routing_mode
    if e2e_device_idx is not None:
        line["e2e_device_idx"] = e2e_device_idx
    if e2e_i32 is not None:
        line["e2e_int32_idx"] = e2e_i32
    if strong is not None:
        line["strong_scaling"] = strong
    if parity is not None:
        line["sharded_parity"] = parity
    if kron is not None:
        line["roofline_kron"] = {"points": kron, "clocks": kron_clocks,
                                 "ncu": "sm__pipe_tensor_cycles_active of the same kernels: profiles/r2_kron_*_full.txt"}
    if world == 1 and not args.no_cpu:
        torch.set_num_threads(host_cores())
        with contextlib.redirect_stdout(sys.stderr):
            sps, t_sample, kind = cpu_crd_steps_per_s(cfg, 4, 1, 256)
            line["cpu_baseline"] = {"value": sps, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": kind,
                                    "sample": f"4 steps of 256 of the 1024 anchors at full K/n ({t_sample * 1e3:.0f} ms each), "
                                              f"scaled x4 (work is linear in anchors)", "extrapolation_factor": 4}
            for which in list(fused):
                try:
                    fused[which]["cpu_baseline"] = (cpu_fused_step("c1", 10, 2) if which == "c1" else
                                                    cpu_fused_step("c4", 2, 1, sample_B=64))
                except Exception as e:               # noqa: BLE001
                    errors["fused_step." + which + ".cpu_baseline"] = f"{type(e).__name__}: {e}"
    if fused:
        line["fused_step"] = fused
    if knn is not None:
        line["knn_positives"] = knn
    if kmeans is not None:
        line["kmeans_centres"] = kmeans
    if errors:
        line["errors"] = errors
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-kron", action="store_true", help="skip the Kronecker / fused-step sections")
    ap.add_argument("--eager", action="store_true", help="time eager kernel launches instead of CUDA-graph replays")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    # stdout carries ONE JSON line.  The modules print Z once (as the reference does) and NCCL prints its version banner
    # from C: python-level prints are collected by _Tee, and file descriptor 1 itself points at stderr while the arm runs.
    buf = _Tee()
    sys.stdout.flush()
    real_fd = os.dup(1)
    os.dup2(2, 1)
    real_stdout, sys.stdout = sys.stdout, buf
    try:
        run_gpu_arm(args)
    finally:
        sys.stdout = real_stdout
        sys.stdout.flush()
        os.dup2(real_fd, 1)
        os.close(real_fd)
    for ln in buf.lines:
        print(ln)
    sys.stdout.flush()


class _Tee:
    """Collects JSON lines printed by the GPU arm; everything else goes to stderr."""

    def __init__(self):
        self.lines = []
        self._part = ""

    def write(self, s):
        self._part += s
        while "\n" in self._part:
            ln, self._part = self._part.split("\n", 1)
            if ln.startswith("{") and ln.endswith("}"):
                self.lines.append(ln)
            elif ln:
                sys.stderr.write(ln + "\n")

    def flush(self):
        pass


if __name__ == "__main__":
    main()
