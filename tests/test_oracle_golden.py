"""Pin the oracle: every oracle function vs the fixtures the UNMODIFIED reference
produced (oracle/make_golden.py).  CPU only."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import crd_oracle as co
from oracle import fusion_oracle as fo

FLOAT_TOL = 2e-6   # oracle and reference run the same torch-CPU ops; only thread/order noise


@pytest.mark.parametrize("name", ["crd_small", "crd_d128", "crd_d64_ragged", "crd_embed1"])
def test_crd_steps_match_reference(golden, name):
    g = golden(name)
    sd = g.state_dict("init.")
    n = g.cfg["n"]
    for s in range(g.cfg["steps"]):
        p = f"step{s}."
        f_s = g.t(p + "f_s").requires_grad_(True)
        f_t = g.t(p + "f_t").requires_grad_(True)
        params = [k for k in sd if k.startswith("embed")]
        for k in params:
            sd[k] = sd[k].detach().requires_grad_(True)
        pre1, pre2 = sd["contrast.memory_v1"].clone(), sd["contrast.memory_v2"].clone()
        z_known = sd["contrast.params"][2].item() > 0
        pre_params = sd["contrast.params"].clone()
        loss, v_s, v_t = co.crd_loss(sd, f_s, f_t, g.t(p + "idx"), g.t(p + "contrast_idx"), n)
        loss.backward()
        assert loss.shape == (1,)
        assert rel_err(loss, g.t(p + "loss")) < FLOAT_TOL
        assert rel_err(f_s.grad, g.t(p + "grad_f_s")) < FLOAT_TOL
        assert rel_err(f_t.grad, g.t(p + "grad_f_t")) < FLOAT_TOL
        for k in params:
            assert rel_err(sd[k].grad, g.t(p + "grad." + k)) < 1e-5, k
        assert rel_err(sd["contrast.params"], g.t(p + "params")) < FLOAT_TOL
        # integer work: exactly the anchors' rows change, nothing else
        for bank, pre in (("memory_v1", pre1), ("memory_v2", pre2)):
            want = g.t(p + bank)
            got = sd["contrast." + bank]
            assert rel_err(got, want) < FLOAT_TOL
            changed = (got != pre).any(dim=1).nonzero().flatten().tolist()
            assert sorted(changed) == sorted(g.t(p + "idx").tolist())
        # closed form (second oracle) agrees with the autograd of the restatement
        if z_known:
            cf_loss, gv1, gv2, x1, x2 = co.crd_closed_form(
                pre1, pre2, v_s.detach(), v_t.detach(), g.t(p + "contrast_idx"),
                pre_params[1].item(), pre_params[2].item(), pre_params[3].item(), n)
            assert rel_err(cf_loss.reshape(1), g.t(p + "loss")) < 1e-5
            assert rel_err(x1.unsqueeze(2), g.t(p + "out_v1")) < 1e-5
            assert rel_err(x2.unsqueeze(2), g.t(p + "out_v2")) < 1e-5


def test_contrast_memory_outputs(golden):
    g = golden("crd_small")
    sd = g.state_dict("init.")
    v_s = co.embed_forward(g.t("step0.f_s"), sd, "embed_s.")
    v_t = co.embed_forward(g.t("step0.f_t"), sd, "embed_t.")
    o1, o2 = co.contrast_memory_forward(sd["contrast.memory_v1"], sd["contrast.memory_v2"],
                                        sd["contrast.params"], v_s, v_t, g.t("step0.idx"),
                                        g.t("step0.contrast_idx"))
    assert o1.shape == (g.cfg["B"], g.cfg["K"] + 1, 1)
    assert rel_err(o1, g.t("step0.out_v1")) < FLOAT_TOL
    assert rel_err(o2, g.t("step0.out_v2")) < FLOAT_TOL


def test_alias_tables_bit_exact(golden):
    g = golden("alias")
    for c in g.cfg["cases"]:
        prob, alias = co.alias_build(g.np(f"{c}.normalised"))
        assert prob.dtype == np.float32 and alias.dtype == np.int64
        assert np.array_equal(prob.view(np.uint32), g.np(f"{c}.prob").view(np.uint32)), c
        assert np.array_equal(alias, g.np(f"{c}.alias")), c
        # normalisation stays a host-side torch call: reproduce it the reference's way
        raw = torch.from_numpy(g.np(f"{c}.raw").copy())
        if raw.sum() > 1:
            raw.div_(raw.sum())
        assert np.array_equal(raw.numpy().view(np.uint32), g.np(f"{c}.normalised").view(np.uint32)), c


def test_alias_draw_select_bit_exact(golden):
    g = golden("alias")
    for c in g.cfg["cases"]:
        out = co.alias_select(g.np(f"{c}.draw_kk"), g.np(f"{c}.draw_b"), g.np(f"{c}.alias"))
        assert np.array_equal(out, g.np(f"{c}.draw_out")), c


def test_alias_uniform_is_identity():
    for n in (1, 2, 1000, 4096, 65536):
        p = torch.ones(n)
        p.div_(p.sum()) if n > 1 else None
        prob, alias = co.alias_build(p.numpy())
        assert (prob == 1).all() and (alias == 0).all()


def _fusion_sd(g):
    sd = g.state_dict("init.")
    for k in sd:
        if sd[k].is_floating_point() and "running" not in k:
            sd[k].requires_grad_(True)
    return sd


def _check_fusion(g, sd, out, ins, tag, tol=2e-5):
    G = g.t(f"{tag}.G")
    (out * G).sum().backward()
    assert rel_err(out, g.t(f"{tag}.out")) < tol
    for i, x in enumerate(ins):
        assert rel_err(x.grad, g.t(f"{tag}.grad_vec{i + 1}")) < tol
    for k in g.keys():
        if k.startswith(f"{tag}.grad.") and not k.startswith(f"{tag}.grad_vec"):
            name = k[len(tag) + 6:]
            got = sd[name].grad if sd[name].grad is not None else torch.zeros_like(sd[name])
            want = g.t(k)
            if want.abs().max() < 1e-4:      # a bias feeding BatchNorm: exact gradient is 0, both hold rounding noise
                assert got.abs().max() < 1e-4, name
            else:
                assert rel_err(got, want) < 5e-5, name


@pytest.mark.parametrize("name", ["bilinear_c1", "bilinear_skip", "bilinear_odd", "bilinear_scaled"])
def test_bilinear_fusion_matches_reference(golden, name):
    g = golden(name)
    kw = {k: g.cfg[k] for k in ("skip", "use_bilinear", "gate1", "gate2") if k in g.cfg}
    modes = ["eval"] + (["train"] if any(k.startswith("train.") for k in g.keys()) else [])
    for tag in modes:
        sd = _fusion_sd(g)
        ins = [g.t("vec1").requires_grad_(True), g.t("vec2").requires_grad_(True)]
        run_sd = {k: (v if v.requires_grad else v.clone()) for k, v in sd.items()}
        out = fo.bilinear_fusion_forward(run_sd, *ins, training=(tag == "train"), **kw)
        _check_fusion(g, sd, out, ins, tag)
        if tag == "train":
            for k in g.keys():
                if k.startswith("train.after."):
                    assert rel_err(run_sd[k[12:]].float(), g.t(k).float()) < 2e-5, k


@pytest.mark.parametrize("name", ["trilinear_A", "trilinear_B"])
def test_trilinear_fusion_matches_reference(golden, name):
    g = golden(name)
    kw = {k: g.cfg[k] for k in ("skip", "use_bilinear", "gate1", "gate2", "gate3") if k in g.cfg}
    sd = _fusion_sd(g)
    ins = [g.t(f"vec{i}").requires_grad_(True) for i in (1, 2, 3)]
    out = fo.trilinear_fusion_forward(sd, *ins, variant=g.cfg["variant"], **kw)
    _check_fusion(g, sd, out, ins, "eval")


def test_kron_linear_is_encoder1(golden):
    g = golden("bilinear_c1")
    torch.manual_seed(0)
    o1, o2 = torch.rand(5, 32), torch.rand(5, 32)
    W, b = g.t("init.encoder1.0.weight"), g.t("init.encoder1.0.bias")
    want = torch.nn.functional.linear(fo.kron_rows(fo._append_one(o1), fo._append_one(o2)), W, b)
    assert rel_err(fo.kron_linear([o1, o2], W, b).float(), want) < 1e-5
    ein = torch.einsum("bi,bj,nij->bn", fo._append_one(o1), fo._append_one(o2), W.view(64, 33, 33)) + b
    assert rel_err(ein, want) < 1e-5


def test_distill_kl(golden):
    g = golden("distill_kl")
    for i in range(g.cfg["cases"]):
        y_s = g.t(f"c{i}.y_s").requires_grad_(True)
        loss = fo.distill_kl(y_s, g.t(f"c{i}.y_t"), float(g.np(f"c{i}.T")))
        loss.backward()
        assert rel_err(loss, g.t(f"c{i}.loss")) < FLOAT_TOL
        assert rel_err(y_s.grad, g.t(f"c{i}.grad_y_s")) < FLOAT_TOL


# ---- selection variant (ContrastMemory_v3 / 5-arg CRDLoss / ContrastLoss_v2), oracle/crd_select_oracle.py ----
SEL_CASES = ["crdsel_random", "crdsel_hard_d128", "crdsel_mid_d64", "crdsel_curriculum", "crdsel_allneg", "crdsel_sampleKD"]


@pytest.mark.parametrize("name", SEL_CASES)
def test_crd_selection_variant_matches_reference(golden, name):
    from oracle import crd_select_oracle as so
    g = golden(name)
    c = g.cfg
    sd = g.state_dict("init.")
    for s in range(c["steps"]):
        p = f"step{s}."
        f_s = g.t(p + "f_s").requires_grad_(True)
        f_t = g.t(p + "f_t").requires_grad_(True)
        params = [k for k in sd if k.startswith("embed")]
        for k in params:
            sd[k] = sd[k].detach().requires_grad_(True)
        pre1, pre2 = sd["contrast.memory_v1"].clone(), sd["contrast.memory_v2"].clone()
        pre_params = sd["contrast.params"].clone()
        np.random.seed(int(g.np(p + "np_seed")))
        loss, out_s, out_t, sel = so.crd_loss_v3(
            sd, float(g.np(p + "epoch")), f_s, f_t, g.t(p + "idx"), g.t(p + "contrast_idx"), c["n"], P2=c["P2"],
            K2=c["K2"], select_pos_mode=c["mode"], select_neg_pairs=c["select_neg_pairs"], sample_KD=c["sample_KD"])
        (loss * g.t(p + "G").reshape(loss.shape)).sum().backward()
        assert rel_err(loss.reshape(-1), g.t(p + "loss")) < FLOAT_TOL
        assert out_s.shape == g.t(p + "out_v1").shape
        assert rel_err(out_s, g.t(p + "out_v1")) < FLOAT_TOL and rel_err(out_t, g.t(p + "out_v2")) < FLOAT_TOL
        assert rel_err(f_s.grad, g.t(p + "grad_f_s")) < FLOAT_TOL
        assert rel_err(f_t.grad, g.t(p + "grad_f_t")) < FLOAT_TOL
        for k in params:
            assert rel_err(sd[k].grad, g.t(p + "grad." + k)) < 1e-5, k
        assert rel_err(sd["contrast.params"], g.t(p + "params")) < FLOAT_TOL
        for bank, pre in (("memory_v1", pre1), ("memory_v2", pre2)):
            got = sd["contrast." + bank]
            assert rel_err(got, g.t(p + bank)) < FLOAT_TOL
            changed = (got != pre).any(dim=1).nonzero().flatten().tolist()
            assert sorted(changed) == sorted(g.t(p + "idx").tolist())
        # integer work: column 0 is the exact positive, positives come from the first P columns, negatives from the rest
        assert (sel[:, 0] == 0).all() and (sel[:, :c["P2"]] < c["P"]).all() and (sel[:, c["P2"]:] >= c["P"]).all()
        # closed form (second oracle; what the fused CUDA kernel evaluates) vs the reference's autograd
        if c["sample_KD"] == "False" and pre_params[2].item() > 0:
            v1 = co.embed_forward(g.t(p + "f_s"), {k: v.detach() for k, v in sd.items()}, "embed_s.")
            v2 = co.embed_forward(g.t(p + "f_t"), {k: v.detach() for k, v in sd.items()}, "embed_t.")
            rows = g.t(p + "contrast_idx").gather(1, sel)
            B, cols = rows.shape
            r1 = pre1.index_select(0, rows.reshape(-1)).view(B, cols, -1)
            r2 = pre2.index_select(0, rows.reshape(-1)).view(B, cols, -1)
            cf, _, _ = so.closed_form_multi_pos(r1, r2, v1, v2, pre_params[1].item(), pre_params[2].item(),
                                                pre_params[3].item(), c["n"], c["P2"])
            assert abs(cf.item() - g.t(p + "loss").item()) < 1e-5 * abs(g.t(p + "loss").item())


@pytest.mark.parametrize("name", ["polynomial_16", "polynomial_gate"])
def test_polynomial_fusion_matches_reference(golden, name):
    """`MIA 2023/stage2_unimodal_student/fusion.py:6-73` (4th-order fusion) -- oracle/make_golden_poly.py."""
    g = golden(name)
    kw = {k: g.cfg[k] for k in ("skip", "use_bilinear", "gate1", "gate2") if k in g.cfg}
    for tag in ("eval", "train"):
        sd = _fusion_sd(g)
        ins = [g.t("vec1").requires_grad_(True), g.t("vec2").requires_grad_(True)]
        run_sd = {k: v.clone() if ("running" in k or "num_batches" in k) else v for k, v in sd.items()}
        out = fo.polynomial_fusion_forward(run_sd, *ins, training=(tag == "train"), **kw)
        _check_fusion(g, sd, out, ins, tag)


# ---- MIA 2022 variants (ContrastMemory_v4 / ContrastMemory_mono, CRD_loss_v2.py), oracle/crd_select_oracle.py ----
V4_CASES = ["crdv4_hard", "crdv4_mid_d128", "crdv4_plain", "crdv4_curriculum_KD"]
MONO_CASES = ["crdmono_hard", "crdmono_mid_d128", "crdmono_random_KD"]


@pytest.mark.parametrize("name", V4_CASES + MONO_CASES)
def test_crd_reweighted_and_mono_variants_match_reference(golden, name):
    from oracle import crd_select_oracle as so
    g = golden(name)
    c = g.cfg
    mono = c["kind"] == "mono"
    sd = g.state_dict("init.")
    for s in range(c["steps"]):
        p = f"step{s}."
        f_s = g.t(p + "f_s").requires_grad_(True)
        f_t = g.t(p + "f_t").requires_grad_(True)
        params = [k for k in sd if k.startswith("embed")]
        for k in params:
            sd[k] = sd[k].detach().requires_grad_(True)
        pre1 = sd["contrast.memory_v1"].clone()
        np.random.seed(int(g.np(p + "np_seed")))
        args = (sd, float(g.np(p + "epoch")), f_s, f_t, g.t(p + "idx"), g.t(p + "contrast_idx"), c["n"])
        if mono:
            loss, out_t, sel_pos = so.crd_loss_mono(*args, P2=c["P2"], select_pos_mode=c["mode"], sample_KD=c["sample_KD"])
        else:
            loss, out_s, out_t, sel_pos = so.crd_loss_v4(*args, P2=c["P2"], select_pos_mode=c["mode"],
                                                         neg_reweight=c["neg_reweight"], sample_KD=c["sample_KD"])
            assert rel_err(out_s, g.t(p + "out_v1")) < FLOAT_TOL
        (loss * g.t(p + "G").reshape(loss.shape)).sum().backward()
        assert out_t.shape == (c["B"], c["P2"] + c["K"], 1)
        assert rel_err(out_t, g.t(p + "out_v2")) < FLOAT_TOL
        assert rel_err(loss.reshape(-1), g.t(p + "loss")) < FLOAT_TOL
        assert rel_err(f_s.grad, g.t(p + "grad_f_s")) < FLOAT_TOL
        if mono:
            assert f_t.grad is None
        else:
            assert rel_err(f_t.grad, g.t(p + "grad_f_t")) < FLOAT_TOL
        for k in params:
            assert rel_err(sd[k].grad, g.t(p + "grad." + k)) < 1e-5, k
        assert rel_err(sd["contrast.params"], g.t(p + "params")) < FLOAT_TOL
        for bank in ("memory_v1", "memory_v2"):
            assert rel_err(sd["contrast." + bank], g.t(p + bank)) < FLOAT_TOL
        changed = (sd["contrast.memory_v1"] != pre1).any(dim=1).nonzero().flatten().tolist()
        assert sorted(changed) == sorted(g.t(p + "idx").tolist())
        assert (sel_pos[:, 0] == 0).all() and (sel_pos < c["P"]).all()


# ---- MIA 2023 stage-2 criterion (CRD_criterion_v10.py: KNN / class-centre positives), oracle/crd_knn_oracle.py ----
KNN_CASES = ["crdknn_p3_d16", "crdknn_p5_d128", "crdknn_p1_d32", "crdknn_centers_d32", "crdknn_kmeans_p4_d32",
             "crdknn_kmeans_p3_d128", "crdknn_kmeans_p6_d64"]


@pytest.mark.parametrize("name", KNN_CASES)
def test_crd_knn_variant_matches_reference(golden, name):
    from oracle import crd_knn_oracle as ko
    g = golden(name)
    c = g.cfg
    sd = g.state_dict("init.")
    cls = g.np("row_class")
    class_idx = [np.nonzero(cls == k)[0] for k in range(3)]
    for s in range(c["steps"]):
        p = f"step{s}."
        f_s = g.t(p + "f_s").requires_grad_(True)
        f_t = g.t(p + "f_t").requires_grad_(True)
        params = [k for k in sd if k.startswith("embed")]
        for k in params:
            sd[k] = sd[k].detach().requires_grad_(True)
        pre1 = sd["contrast.memory_v1"].clone()
        kmeans = c["pos_extra"] == "centers" and c["P"] > 2
        if kmeans:                                          # the k-means restatement against sklearn's fits inside the reference
            for b, bank in enumerate(("memory_v1", "memory_v2")):
                for k in range(3):
                    X = sd["contrast." + bank][torch.as_tensor(class_idx[k])].numpy()
                    got, iters = ko.kmeans_lloyd(X, g.np(p + "kmeans_init")[b, k])
                    assert iters == int(g.np(p + "kmeans_iters")[3 * b + k])
                    assert np.abs(got - g.np(p + "kmeans_centres")[b, k]).max() < 2e-6
        loss, sample_loss, res = ko.crd_loss_v10(sd, class_idx, c["P"], c["pos_extra"], g.t(p + "sample_weights"), f_s, f_t,
                                                 g.t(p + "label"), g.t(p + "idx"), g.t(p + "contrast_idx"), c["n"],
                                                 kmeans_init=g.t(p + "kmeans_init") if kmeans else None)
        loss.backward()
        width = c["P"] + c["K"] if c["pos_extra"] == "neighbors" else (c["P"] - 1) + (c["K"] + 1) + 2 * (c["P"] - 1)
        assert res[0].shape == (c["B"], width, 1)
        assert rel_err(res[0], g.t(p + "out_v1")) < FLOAT_TOL and rel_err(res[1], g.t(p + "out_v2")) < FLOAT_TOL
        if c["pos_extra"] == "neighbors":
            assert rel_err(res[2], g.t(p + "sim_v1")) < FLOAT_TOL and rel_err(res[3], g.t(p + "sim_v2")) < FLOAT_TOL
            assert (res[2][:, 0] - 1).abs().max() < 1e-5          # the first neighbour is the query itself (:108)
        assert rel_err(loss.reshape(-1), g.t(p + "loss")) < FLOAT_TOL
        assert rel_err(sample_loss, g.t(p + "sample_loss")) < FLOAT_TOL
        assert rel_err(f_s.grad, g.t(p + "grad_f_s")) < 1e-5
        assert rel_err(f_t.grad, g.t(p + "grad_f_t")) < 1e-5
        for k in params:
            assert rel_err(sd[k].grad, g.t(p + "grad." + k)) < 1e-5, k
        assert rel_err(sd["contrast.params"], g.t(p + "params")) < FLOAT_TOL
        for bank in ("memory_v1", "memory_v2"):
            assert rel_err(sd["contrast." + bank], g.t(p + bank)) < FLOAT_TOL
        changed = (sd["contrast.memory_v1"] != pre1).any(dim=1).nonzero().flatten().tolist()
        assert sorted(changed) == sorted(g.t(p + "idx").tolist())
