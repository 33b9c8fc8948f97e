"""GPU parity tests of the selection variant of CRD (reference: CL_utils/CRD_loss.py:127-175 5-arg CRDLoss over
CL_utils/memory_new.py:225-397 ContrastMemory_v3 and CRD_loss.py:212-252 ContrastLoss_v2): the CUDA path (through the
C ABI and the drop-in modules) vs the reference-generated goldens and the CPU oracle.

Tolerances: selected columns / touched rows bit-exact; floats rel 1e-4 (fp32 memory path)."""
from __future__ import annotations

import types

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"
CASES = ["crdsel_random", "crdsel_hard_d128", "crdsel_mid_d64", "crdsel_curriculum", "crdsel_allneg", "crdsel_sampleKD"]


@pytest.fixture(scope="module")
def pkg():
    import multimodal_learning_b200 as p
    assert torch.cuda.is_available()
    p._cabi.lib()
    return p


@pytest.fixture(scope="module")
def so():
    from oracle import crd_select_oracle
    return crd_select_oracle


def _opt(c):
    return types.SimpleNamespace(s_dim=c["s_dim"], t_dim=c["t_dim"], feat_dim=c["D"], nce_p=c["P"], nce_p2=c["P2"],
                                 nce_k=c["K"], nce_k2=c["K2"], nce_t=c["T"], nce_m=c["momentum"], select_pos_pairs=True,
                                 select_neg_pairs=c["select_neg_pairs"], sample_KD=c["sample_KD"], select_pos_mode=c["mode"])


def _module(pkg, g):
    mod = pkg.crd_select.CRDLoss(_opt(g.cfg), g.cfg["n"])
    mod.load_state_dict(g.state_dict("init."))
    return mod.to(DEV)


@pytest.mark.parametrize("name", CASES)
def test_crdloss_5arg_steps_match_reference_golden(pkg, golden, name, capsys):
    g = golden(name)
    c = g.cfg
    mod = _module(pkg, g)
    before = pkg._cabi.launch_count()
    for s in range(c["steps"]):
        p = f"step{s}."
        f_s = g.t(p + "f_s", DEV).requires_grad_(True)
        f_t = g.t(p + "f_t", DEV).requires_grad_(True)
        idx, cidx = g.t(p + "idx", DEV), g.t(p + "contrast_idx", DEV)
        pre1 = mod.contrast.memory_v1.clone()
        mod.zero_grad()
        np.random.seed(int(g.np(p + "np_seed")))          # the reference's picks come from the global numpy RNG
        loss = mod(float(g.np(p + "epoch")), f_s, f_t, idx, cidx)
        assert loss.shape == (() if c["sample_KD"] == "False" else (c["B"],))
        (loss * g.t(p + "G", DEV).reshape(loss.shape)).sum().backward()
        assert rel_err(loss.reshape(-1), g.t(p + "loss")) < TOL
        assert rel_err(f_s.grad, g.t(p + "grad_f_s")) < TOL
        assert rel_err(f_t.grad, g.t(p + "grad_f_t")) < TOL
        for k, v in mod.named_parameters():
            assert rel_err(v.grad, g.t(p + "grad." + k)) < TOL, k
        assert rel_err(mod.contrast.params, g.t(p + "params")) < TOL
        assert rel_err(mod.contrast.memory_v1, g.t(p + "memory_v1")) < TOL
        assert rel_err(mod.contrast.memory_v2, g.t(p + "memory_v2")) < TOL
        changed = (mod.contrast.memory_v1 != pre1).any(dim=1).nonzero().flatten().cpu().tolist()
        assert sorted(changed) == sorted(g.t(p + "idx").tolist())
    assert pkg._cabi.launch_count() > before
    assert "normalization constant Z_v1 is set to" in capsys.readouterr().out


@pytest.mark.parametrize("name", ["crdsel_random", "crdsel_hard_d128", "crdsel_mid_d64", "crdsel_allneg"])
def test_contrast_memory_v3_direct_outputs_selection_and_autograd(pkg, so, golden, name):
    """ContrastMemory_v3.forward returns the SELECTED, Z-normalised scores [B, P2+K2, 1] in the reference's column
    order; the selected columns are bit-exact vs the oracle's selection; gradients use the pre-update rows."""
    from oracle import crd_oracle as co
    g = golden(name)
    c = g.cfg
    sd = g.state_dict("init.")
    mem = pkg.ContrastMemory_v3(c["D"], c["n"], c["P"], c["K"], c["T"], c["momentum"], True, c["P2"],
                                c["select_neg_pairs"], c["K2"])
    mem.load_state_dict({k[len("contrast."):]: v for k, v in sd.items() if k.startswith("contrast.")})
    mem = mem.to(DEV)
    v_s = co.embed_forward(g.t("step0.f_s"), sd, "embed_s.")
    v_t = co.embed_forward(g.t("step0.f_t"), sd, "embed_t.")
    idx, cidx = g.t("step0.idx"), g.t("step0.contrast_idx")
    epoch = float(g.np("step0.epoch"))
    # integer work: the kernel's relation gap orders the columns exactly like the oracle's
    np.random.seed(int(g.np("step0.np_seed")))
    sel, sel_idx = mem.select(epoch, v_s.to(DEV), v_t.to(DEV), cidx.to(DEV), c["mode"])
    t_rel, s_rel = so.relations(sd["contrast.memory_v1"], sd["contrast.memory_v2"], v_s, v_t, cidx)
    diff = (t_rel - s_rel).squeeze(-1)
    got_diff = pkg.crd_select.crd_relation_diff(mem.memory_v1, mem.memory_v2, v_s.to(DEV), v_t.to(DEV), cidx.to(DEV))
    assert rel_err(got_diff, diff) < TOL
    np.random.seed(int(g.np("step0.np_seed")))
    picks = so.positive_picks(c["mode"], epoch, c["P"], c["P2"])
    want_sel = so.select_columns(diff, c["P"], c["K"], c["P2"], c["K2"], picks, c["select_neg_pairs"])
    # columns that sample the SAME bank row tie exactly in diff and torch.sort / topk may order them either way (also
    # inside the reference, whose CUDA sort is not stable): compare the selected ROWS, which is what the loss sees
    assert torch.equal(sel_idx.cpu(), cidx.gather(1, want_sel))
    assert torch.equal(sel_idx.cpu(), cidx.gather(1, sel.cpu()))
    assert (sel[:, 0] == 0).all() and (sel[:, :c["P2"]] < c["P"]).all() and (sel[:, c["P2"]:] >= c["P"]).all()
    # forward + autograd through the selected scores
    a = v_s.clone().requires_grad_(True)
    b = v_t.clone().requires_grad_(True)
    np.random.seed(int(g.np("step0.np_seed")))
    o1, o2, _ = so.contrast_memory_v3_forward(sd["contrast.memory_v1"].clone(), sd["contrast.memory_v2"].clone(),
                                              sd["contrast.params"].clone(), epoch, a, b, idx, cidx, P2=c["P2"], K2=c["K2"],
                                              select_pos_mode=c["mode"], select_neg_pairs=c["select_neg_pairs"])
    torch.manual_seed(3)
    G1, G2 = torch.rand_like(o1), torch.rand_like(o2)
    ((o1 * G1).sum() + (o2 * G2).sum()).backward()
    ad = v_s.to(DEV).requires_grad_(True)
    bd = v_t.to(DEV).requires_grad_(True)
    np.random.seed(int(g.np("step0.np_seed")))
    d1, d2 = mem(epoch, ad, bd, idx.to(DEV), cidx.to(DEV), c["mode"])
    assert d1.shape == g.t("step0.out_v1").shape
    assert rel_err(d1, g.t("step0.out_v1")) < TOL and rel_err(d2, g.t("step0.out_v2")) < TOL
    ((d1 * G1.to(DEV)).sum() + (d2 * G2.to(DEV)).sum()).backward()
    assert rel_err(ad.grad, a.grad) < TOL and rel_err(bd.grad, b.grad) < TOL


@pytest.mark.parametrize("B,D,P,K,P2,K2,n", [
    (16, 128, 300, 700, 10, 512, 1024),       # the reference's defaults (options.py:85-91)
    (3, 64, 5, 9, 2, 4, 40),
    (2, 48, 4, 6, 1, 3, 20),                  # generic-D kernels, single positive
    (4, 256, 6, 40, 3, 40, 64),               # K2 == K
])
def test_multipos_kernel_vs_closed_form(pkg, so, B, D, P, K, P2, K2, n):
    gen = torch.Generator().manual_seed(B * 31 + D)
    stdv = 1.0 / (D / 3) ** 0.5
    m1 = torch.rand(n, D, generator=gen) * 2 * stdv - stdv
    m2 = torch.rand(n, D, generator=gen) * 2 * stdv - stdv
    v1 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=1)
    v2 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=1)
    sel_idx = torch.randint(0, n, (B, P2 + K2), generator=gen)
    T = float(torch.tensor(0.07).item())
    Z = torch.tensor([37.0, 41.0])
    d = lambda t: t.to(DEV)
    loss, g1, g2, o1, o2 = pkg.crd_select.crd_fused_loss_grad_multipos(d(m1), d(m2), d(v1), d(v2), d(sel_idx), P2, T, d(Z),
                                                                       n, want_out=True)
    r1 = m1.index_select(0, sel_idx.reshape(-1)).view(B, P2 + K2, D)
    r2 = m2.index_select(0, sel_idx.reshape(-1)).view(B, P2 + K2, D)
    w_loss, w_g1, w_g2 = so.closed_form_multi_pos(r1, r2, v1, v2, T, 37.0, 41.0, n, P2)
    assert abs(loss.item() - w_loss.item()) < TOL * abs(w_loss.item())
    assert rel_err(g1, w_g1) < TOL and rel_err(g2, w_g2) < TOL
    x1 = torch.exp(torch.einsum("bkd,bd->bk", r2.double(), v1.double()) / T) / 37.0
    assert rel_err(o1, x1) < TOL
    # and against the stand-alone criterion module fed with the kernel's own scores
    crit = pkg.ContrastLoss_v2(n, "False")
    ref = crit(o1.unsqueeze(2), P2) + crit(o2.unsqueeze(2), P2)
    assert abs(loss.item() - ref.item()) < TOL * abs(ref.item())
    if P2 == 1:     # n_pos = 1 is the plain CRD criterion, bit for bit
        from multimodal_learning_b200 import crd
        l1, _, h1, h2, _, _ = crd.crd_fused_loss_grad(d(m1), d(m2), d(v1), d(v2), d(sel_idx), T, d(Z), n, K2)
        assert torch.equal(l1, loss) and torch.equal(h1, g1) and torch.equal(h2, g2)


def test_sampled_idx_and_errors(pkg, golden):
    g = golden("crdsel_random")
    c = g.cfg
    mod = _module(pkg, g)
    f_s, f_t, idx = g.t("step0.f_s", DEV), g.t("step0.f_t", DEV), g.t("step0.idx", DEV)
    np.random.seed(0)
    loss = mod(0.0, f_s, f_t, idx, None)              # contrast_idx drawn on the device (memory_new.py:263-265)
    assert torch.isfinite(loss).all()
    bad = g.t("step0.contrast_idx", DEV)[:, :-1].contiguous()
    with pytest.raises(RuntimeError):
        mod(0.0, f_s, f_t, idx, bad)
    with pytest.raises(RuntimeError):                 # CPU tensors: no fallback
        mod(0.0, g.t("step0.f_s"), g.t("step0.f_t"), g.t("step0.idx"), g.t("step0.contrast_idx"))


def test_weighted_crdloss_matches_oracle(pkg, golden):
    """weighted_CRDLoss (CRD_loss.py:8-50) = plain ContrastMemory + per-sample gated criterion."""
    from oracle import crd_oracle as co
    g = golden("crd_small")
    c = g.cfg
    opt = types.SimpleNamespace(s_dim=c["s_dim"], t_dim=c["t_dim"], feat_dim=c["D"], nce_k=c["K"], nce_t=0.07, nce_m=0.5)
    mod = pkg.crd_select.weighted_CRDLoss(opt, c["n"])
    sd = g.state_dict("init.")
    # single-Linear heads in this file (CRD_loss.py:256-267): take the first Linear of the golden's 2-layer head
    own = mod.state_dict()
    for k in own:
        src = k.replace("linear.", "linear.0.") if k.startswith("embed") else k
        own[k] = sd[src].clone()
    mod.load_state_dict(own)
    mod = mod.to(DEV)
    f_s, f_t = g.t("step0.f_s"), g.t("step0.f_t")
    idx, cidx = g.t("step0.idx"), g.t("step0.contrast_idx")
    torch.manual_seed(1)
    ls, lt = torch.rand(c["B"], 1), torch.rand(c["B"], 1)
    loss = mod(f_s.to(DEV), f_t.to(DEV), ls.to(DEV), lt.to(DEV), idx.to(DEV), cidx.to(DEV))
    sd1 = {"embed_s.linear.weight": own["embed_s.linear.weight"], "embed_s.linear.bias": own["embed_s.linear.bias"],
           "embed_t.linear.weight": own["embed_t.linear.weight"], "embed_t.linear.bias": own["embed_t.linear.bias"]}
    v1, v2 = co.embed_forward(f_s, sd1, "embed_s."), co.embed_forward(f_t, sd1, "embed_t.")
    o1, o2 = co.contrast_memory_forward(sd["contrast.memory_v1"].clone(), sd["contrast.memory_v2"].clone(),
                                        sd["contrast.params"].clone(), v1, v2, idx, cidx)

    def crit(x, w):            # CRD_loss.py:62-81
        m = x.size(1) - 1
        Pn = 1 / float(c["n"])
        pos = x.select(1, 0)
        log_D1 = torch.div(pos, pos.add(m * Pn + 1e-7)).log()
        neg = x.narrow(1, 1, m)
        log_D0 = torch.div(neg.clone().fill_(m * Pn), neg.add(m * Pn + 1e-7)).log()
        return -torch.sum(w * (log_D1 + log_D0.view(x.shape[0], -1).sum(1, keepdims=True))) / x.shape[0]
    want = crit(o1, torch.where(ls > lt, 1.0, 0.0)) + crit(o2, torch.where(lt > ls, 1.0, 0.0))
    assert abs(loss.item() - want.item()) < TOL * abs(want.item())


def test_graphed_train_step_matches_eager(pkg, golden):
    """GraphedTrainStep (one CUDA-graph launch per step) == the eager step, bit for bit, over several steps with
    different inputs, including the bank updates and the Adam state."""
    g = golden("crd_d128")
    c = g.cfg
    opt = types.SimpleNamespace(s_dim=c["s_dim"], t_dim=c["t_dim"], feat_dim=c["D"], n_data=c["n"], nce_k=c["K"], nce_t=0.07,
                                nce_m=0.5)
    gen = torch.Generator(device=DEV).manual_seed(5)

    def inputs():
        f_s = torch.randn(c["B"], c["s_dim"], device=DEV, generator=gen)
        f_t = torch.randn(c["B"], c["t_dim"], device=DEV, generator=gen)
        idx = torch.randperm(c["n"], device=DEV, generator=gen)[:c["B"]].contiguous()
        cidx = torch.randint(0, c["n"], (c["B"], c["K"] + 1), device=DEV, generator=gen)
        cidx[:, 0] = idx
        return f_s, f_t, idx, cidx
    batches = [inputs() for _ in range(6)]
    results = []
    for mode in ("eager", "graph"):
        mod = pkg.CRDLoss(opt)
        mod.load_state_dict(g.state_dict("init."))
        mod = mod.to(DEV)
        params = list(mod.parameters())
        optim = torch.optim.Adam(params, lr=1e-3, capturable=True, fused=True)
        losses, grads = [], []
        if mode == "graph":
            step = pkg.GraphedTrainStep(lambda a, b, i, ci: mod(a, b, i, ci), params, optim, batches[0], grad_inputs=(0,),
                                        warmup=0 + 1, n_buffers=2)
            # the constructor ran 1 eager warm-up step on batches[0]; mirror that in the eager arm below
            for b in batches[1:]:
                losses.append(step(*b).clone())
                grads.append(step.static_in[(step._next - 1) % 2][0].grad.clone())
        else:
            for i, b in enumerate(batches):
                f_s = b[0].clone().requires_grad_(True)
                for p in params:
                    p.grad = None
                loss = mod(f_s, *b[1:])
                loss.backward()
                optim.step()
                if i > 0:
                    losses.append(loss.detach().clone())
                    grads.append(f_s.grad.clone())
        results.append((losses, grads, mod.contrast.memory_v1.clone(), [p.detach().clone() for p in params]))
    (l0, g0, m0, p0), (l1, g1, m1, p1) = results
    for a, b in zip(l0, l1):
        assert rel_err(b, a) < 1e-6
    for a, b in zip(g0, g1):
        assert rel_err(b, a) < 1e-5
    assert rel_err(m1, m0) < 1e-6
    for a, b in zip(p0, p1):
        assert rel_err(b, a) < 1e-5


@pytest.mark.parametrize("B,P,K,K2", [(7, 300, 700, 256), (5, 40, 200, 64), (3, 1, 33, 33), (16, 150, 16384, 8192)])
def test_sort_columns_kernel_matches_library_sort(pkg, B, P, K, K2):
    """`mml_crd_sort_columns` (per-anchor bitonic network in shared memory) against torch.sort on the same gaps: the
    descending order of the P positive columns (memory_new.py:303) and the K2 smallest of the K negative columns in
    ascending order (:342-345).  Distinct values: identical column numbers; exact ties: ordered by column."""
    from multimodal_learning_b200.crd_select import sort_columns
    gen = torch.Generator(device=DEV).manual_seed(B + K)
    diff = torch.randn(B, P + K, device=DEV, generator=gen)
    before = pkg._cabi.launch_count()
    pos = sort_columns(diff, 0, P, descending=True)
    neg = sort_columns(diff, P, K, descending=False, first=K2, label0=P)
    assert pkg._cabi.launch_count() == before + 2
    assert torch.equal(pos, torch.sort(diff[:, :P], dim=1, descending=True, stable=True)[1])
    assert torch.equal(neg, P + torch.sort(diff[:, P:], dim=1, stable=True)[1][:, :K2])
    # ties (the same bank row sampled twice gives bit-identical gaps): ordered by column number, values still sorted
    tied = diff.clone()
    tied[:, 3] = tied[:, 0]
    tied[:, P + 5] = tied[:, P + 1]
    tied[:, P + 9] = tied[:, P + 1]
    pos_t = sort_columns(tied, 0, P, descending=True)
    neg_t = sort_columns(tied, P, K, descending=False, first=K2, label0=P)
    assert torch.equal(pos_t, torch.sort(tied[:, :P], dim=1, descending=True, stable=True)[1])
    assert torch.equal(neg_t, P + torch.sort(tied[:, P:], dim=1, stable=True)[1][:, :K2])


# ---- MIA 2022 variants: CRD_loss_v2.py CRDLoss (ContrastMemory_v4) and CRDLoss_v2 (ContrastMemory_mono) ----
V4_CASES = ["crdv4_hard", "crdv4_mid_d128", "crdv4_plain", "crdv4_curriculum_KD"]
MONO_CASES = ["crdmono_hard", "crdmono_mid_d128", "crdmono_random_KD"]


def _opt_v4(c):
    return types.SimpleNamespace(s_dim=c["s_dim"], t_dim=c["t_dim"], feat_dim=c["D"], nce_p=c["P"], nce_p2=c["P2"],
                                 nce_k=c["K"], nce_k2=c["K"], nce_t=c["T"], nce_m=c["momentum"], select_pos_pairs=True,
                                 select_neg_pairs="False", neg_reweight=c["neg_reweight"], sample_KD=c["sample_KD"],
                                 select_pos_mode=c["mode"])


@pytest.mark.parametrize("name", V4_CASES + MONO_CASES)
def test_crd_loss_v2_modules_match_reference_golden(pkg, golden, name, capsys):
    """Reference fixtures from oracle/make_golden_select_v4.py (unmodified `MIA 2022/CL_utils/CRD_loss_v2.py`)."""
    g = golden(name)
    c = g.cfg
    mono = c["kind"] == "mono"
    cls = pkg.crd_loss_v2.CRDLoss_v2 if mono else pkg.crd_loss_v2.CRDLoss
    mod = cls(_opt_v4(c), c["n"])
    mod.load_state_dict(g.state_dict("init."))
    mod = mod.to(DEV)
    captured = {}
    mod.contrast.register_forward_hook(lambda m, i, o: captured.update(out=o))
    before = pkg._cabi.launch_count()
    for s in range(c["steps"]):
        p = f"step{s}."
        f_s = g.t(p + "f_s", DEV).requires_grad_(True)
        f_t = g.t(p + "f_t", DEV).requires_grad_(True)
        idx, cidx = g.t(p + "idx", DEV), g.t(p + "contrast_idx", DEV)
        pre1 = mod.contrast.memory_v1.clone()
        mod.zero_grad()
        np.random.seed(int(g.np(p + "np_seed")))
        loss = mod(float(g.np(p + "epoch")), f_s, f_t, idx, cidx)
        assert loss.shape == (() if c["sample_KD"] == "False" else (c["B"],))
        (loss * g.t(p + "G", DEV).reshape(loss.shape)).sum().backward()
        if mono:
            out_t, bank = captured["out"]
            assert bank is mod.contrast.memory_v1 and mod.memory_t is bank            # CRD_loss_v2.py:100
            assert f_t.grad is None
        else:
            out_s, out_t = captured["out"]
            assert rel_err(out_s, g.t(p + "out_v1")) < TOL
            assert rel_err(f_t.grad, g.t(p + "grad_f_t")) < TOL
        assert out_t.shape == (c["B"], c["P2"] + c["K"], 1) and out_t.is_contiguous()
        assert rel_err(out_t, g.t(p + "out_v2")) < TOL
        assert rel_err(loss.reshape(-1), g.t(p + "loss")) < TOL
        assert rel_err(f_s.grad, g.t(p + "grad_f_s")) < TOL
        for k, v in mod.named_parameters():
            assert rel_err(v.grad, g.t(p + "grad." + k)) < TOL, k
        assert rel_err(mod.contrast.params, g.t(p + "params")) < TOL
        assert rel_err(mod.contrast.memory_v1, g.t(p + "memory_v1")) < TOL
        assert rel_err(mod.contrast.memory_v2, g.t(p + "memory_v2")) < TOL
        changed = (mod.contrast.memory_v1 != pre1).any(dim=1).nonzero().flatten().cpu().tolist()
        assert sorted(changed) == sorted(g.t(p + "idx").tolist())
    assert pkg._cabi.launch_count() > before
    assert "normalization constant Z_v2 is set to" in capsys.readouterr().out


@pytest.mark.parametrize("name", ["crdv4_hard", "crdmono_hard"])
def test_v4_and_mono_no_grad_path_and_state_dict_round_trip(pkg, so, golden, name):
    """Evaluation (no autograd) goes through the scores kernel directly; a reloaded state dict keeps Z and the layout."""
    g = golden(name)
    c = g.cfg
    mono = c["kind"] == "mono"
    cls = pkg.ContrastMemory_mono if mono else pkg.ContrastMemory_v4
    mem = cls(c["D"], c["n"], c["P"], c["K"], c["T"], c["momentum"], True, c["P2"], "False", c["neg_reweight"], c["K"])
    sd = {k[len("contrast."):]: v for k, v in g.state_dict("init.").items() if k.startswith("contrast.")}
    mem.load_state_dict(sd)
    mem = mem.to(DEV)
    gen = torch.Generator().manual_seed(5)
    for step in range(2):
        v1 = torch.nn.functional.normalize(torch.randn(c["B"], c["D"], generator=gen))
        v2 = torch.nn.functional.normalize(torch.randn(c["B"], c["D"], generator=gen))
        y = torch.randperm(c["n"], generator=gen)[:c["B"]]
        idx = torch.randint(0, c["n"], (c["B"], c["P"] + c["K"]), generator=gen)
        idx[:, 0] = y
        ref = {k: v.clone() for k, v in sd.items()}
        if mono:
            want, _ = so.contrast_memory_mono_forward(ref["memory_v1"], ref["memory_v2"], ref["params"], 0.0, v1, v2, y, idx,
                                                      P2=c["P2"], select_pos_mode="hard")
        else:
            _, want, _ = so.contrast_memory_v4_forward(ref["memory_v1"], ref["memory_v2"], ref["params"], 0.0, v1, v2, y,
                                                       idx, P2=c["P2"], select_pos_mode="hard",
                                                       neg_reweight=c["neg_reweight"])
        with torch.no_grad():
            got = mem(0.0, v1.to(DEV), v2.to(DEV), y.to(DEV), idx.to(DEV), "hard")
        got = got[0] if mono else got[1]
        assert rel_err(got, want) < TOL
        assert rel_err(mem.params, ref["params"]) < TOL and rel_err(mem.memory_v1, ref["memory_v1"]) < TOL
        sd = ref
        clone = cls(c["D"], c["n"], c["P"], c["K"], c["T"], c["momentum"], True, c["P2"], "False", c["neg_reweight"], c["K"])
        clone.load_state_dict(mem.state_dict())
        assert clone._z_ready and clone._P == c["P"] and clone._K == c["K"]
