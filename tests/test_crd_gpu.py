"""GPU parity tests of the CRD path: the CUDA kernels (through the C ABI and the drop-in
modules) vs the CPU oracle and the reference-generated golden fixtures.

Tolerances (north_star): integer work bit-exact; memory path rel 1e-4 (fp32 accumulate).
`rel` is max|a-b| / max|b| (conftest.rel_err)."""
from __future__ import annotations

import types

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def pkg():
    import multimodal_learning_b200 as p
    assert torch.cuda.is_available()
    p._cabi.lib()
    return p


@pytest.fixture(scope="module")
def co():
    from oracle import crd_oracle
    return crd_oracle


DEV = "cuda:0"


def _make_module(pkg, cfg, sd, single_linear=False):
    opt = types.SimpleNamespace(s_dim=cfg["s_dim"], t_dim=cfg["t_dim"], feat_dim=cfg["D"], n_data=cfg["n"],
                                nce_k=cfg["K"], nce_t=cfg.get("T", 0.07), nce_m=cfg.get("momentum", 0.5))
    mod = pkg.CRDLoss(opt)
    if single_linear:
        mod.embed_s = pkg.Embed(cfg["s_dim"], cfg["D"], layers=1)
        mod.embed_t = pkg.Embed(cfg["t_dim"], cfg["D"], layers=1)
    mod.load_state_dict(sd)
    return mod.to(DEV)


@pytest.mark.parametrize("name", ["crd_small", "crd_d128", "crd_d64_ragged", "crd_embed1"])
def test_crdloss_steps_match_reference_golden(pkg, golden, name, capsys):
    g = golden(name)
    mod = _make_module(pkg, g.cfg, g.state_dict("init."), single_linear=(name == "crd_embed1"))
    before = pkg._cabi.launch_count()
    for s in range(g.cfg["steps"]):
        p = f"step{s}."
        f_s = g.t(p + "f_s", DEV).requires_grad_(True)
        f_t = g.t(p + "f_t", DEV).requires_grad_(True)
        idx, cidx = g.t(p + "idx", DEV), g.t(p + "contrast_idx", DEV)
        cidx_copy = cidx.clone()
        pre1 = mod.contrast.memory_v1.clone()
        mod.zero_grad()
        loss = mod(f_s, f_t, idx, cidx)
        loss.backward()
        assert loss.shape == (1,)
        assert rel_err(loss, g.t(p + "loss")) < TOL
        assert rel_err(f_s.grad, g.t(p + "grad_f_s")) < TOL
        assert rel_err(f_t.grad, g.t(p + "grad_f_t")) < TOL
        for k, v in mod.named_parameters():
            assert rel_err(v.grad, g.t(p + "grad." + k)) < TOL, k
        assert rel_err(mod.contrast.params, g.t(p + "params")) < TOL          # Z_v1, Z_v2 set on step 0
        assert rel_err(mod.contrast.memory_v1, g.t(p + "memory_v1")) < TOL
        assert rel_err(mod.contrast.memory_v2, g.t(p + "memory_v2")) < TOL
        # integer work, bit exact: only the anchors' rows changed; the caller's idx is untouched
        changed = (mod.contrast.memory_v1 != pre1).any(dim=1).nonzero().flatten().cpu().tolist()
        assert sorted(changed) == sorted(g.t(p + "idx").tolist())
        assert torch.equal(cidx, cidx_copy)
    assert pkg._cabi.launch_count() > before
    assert "normalization constant Z_v1 is set to" in capsys.readouterr().out     # reference side effect (:55)


def test_contrast_memory_direct_outputs_and_autograd(pkg, co, golden):
    """ContrastMemory.forward returns (out_v1, out_v2) [B,K+1,1] with gradients to v1/v2 that use the
    PRE-update rows, even though the banks were updated inside forward (reference autograd semantics)."""
    g = golden("crd_small")
    c = g.cfg
    sd = g.state_dict("init.")
    mem = pkg.ContrastMemory(c["D"], c["n"], c["K"])
    mem.load_state_dict({k[len("contrast."):]: v for k, v in sd.items() if k.startswith("contrast.")})
    mem = mem.to(DEV)
    v_s = co.embed_forward(g.t("step0.f_s"), sd, "embed_s.")
    v_t = co.embed_forward(g.t("step0.f_t"), sd, "embed_t.")
    idx, cidx = g.t("step0.idx"), g.t("step0.contrast_idx")
    # oracle with autograd
    a = v_s.clone().requires_grad_(True)
    b = v_t.clone().requires_grad_(True)
    o1, o2 = co.contrast_memory_forward(sd["contrast.memory_v1"], sd["contrast.memory_v2"], sd["contrast.params"],
                                        a, b, idx, cidx)
    torch.manual_seed(3)
    G1, G2 = torch.rand_like(o1), torch.rand_like(o2)
    ((o1 * G1).sum() + (o2 * G2).sum()).backward()
    # CUDA
    ad = v_s.to(DEV).requires_grad_(True)
    bd = v_t.to(DEV).requires_grad_(True)
    d1, d2 = mem(ad, bd, idx.to(DEV), cidx.to(DEV))
    assert d1.shape == (c["B"], c["K"] + 1, 1) and d2.shape == d1.shape
    assert rel_err(d1, g.t("step0.out_v1")) < TOL and rel_err(d2, g.t("step0.out_v2")) < TOL
    # a second forward (another update of partly the same rows) BEFORE backward must not disturb it
    with torch.no_grad():
        mem(ad.detach(), bd.detach(), idx.to(DEV), cidx.to(DEV))
    ((d1 * G1.to(DEV)).sum() + (d2 * G2.to(DEV)).sum()).backward()
    assert rel_err(ad.grad, a.grad) < TOL
    assert rel_err(bd.grad, b.grad) < TOL


def test_backward_is_repeatable_with_retain_graph(pkg, golden):
    """train_test_path_multi_distill.py:49-56 calls backward(retain_graph=True) several times."""
    g = golden("crd_small")
    mod = _make_module(pkg, g.cfg, g.state_dict("init."))
    f_s = g.t("step0.f_s", DEV).requires_grad_(True)
    loss = mod(f_s, g.t("step0.f_t", DEV), g.t("step0.idx", DEV), g.t("step0.contrast_idx", DEV))
    g1, = torch.autograd.grad(loss, f_s, retain_graph=True)
    g2, = torch.autograd.grad(loss, f_s, retain_graph=True)
    assert torch.equal(g1, g2)
    assert rel_err(g1, g.t("step0.grad_f_s")) < TOL


def test_int32_contrast_idx_gives_identical_bits(pkg, golden):
    """Extension: the caller may hand contrast_idx over as int32 (half the H2D bytes); same kernels, same bits."""
    g = golden("crd_d128")
    res = []
    for dt in (torch.int64, torch.int32):
        mod = _make_module(pkg, g.cfg, g.state_dict("init."))
        f_s = g.t("step0.f_s", DEV).requires_grad_(True)
        loss = mod(f_s, g.t("step0.f_t", DEV), g.t("step0.idx", DEV), g.t("step0.contrast_idx", DEV).to(dt))
        loss.backward()
        res.append((loss.detach().clone(), f_s.grad.clone(), mod.contrast.memory_v1.clone(), mod.contrast.params.clone()))
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_wrong_idx_width_raises_like_reference(pkg, golden):
    g = golden("crd_small")
    mod = _make_module(pkg, g.cfg, g.state_dict("init."))
    bad = g.t("step0.contrast_idx", DEV)[:, :-1].contiguous()
    with pytest.raises(RuntimeError):
        mod(g.t("step0.f_s", DEV), g.t("step0.f_t", DEV), g.t("step0.idx", DEV), bad)
    with pytest.raises(RuntimeError):      # CPU tensors: no fallback
        mod(g.t("step0.f_s"), g.t("step0.f_t"), g.t("step0.idx"), g.t("step0.contrast_idx"))


def _random_problem(B, D, K, n, seed, unit_rows=True):
    gen = torch.Generator().manual_seed(seed)
    stdv = 1.0 / (D / 3) ** 0.5
    m1 = torch.rand(n, D, generator=gen) * 2 * stdv - stdv
    m2 = torch.rand(n, D, generator=gen) * 2 * stdv - stdv
    v1 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=1)
    v2 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=1)
    y = torch.randperm(n, generator=gen)[:B]
    idx = torch.randint(0, n, (B, K + 1), generator=gen)
    idx[:, 0] = y
    return m1, m2, v1, v2, y, idx


@pytest.mark.parametrize("B,D,K,n", [
    (64, 128, 4096, 4096),      # BASELINE config 1 (CRD part)
    (3, 128, 7, 50),            # fewer columns than one warp block
    (5, 64, 100, 77),           # D=64 fast path, ragged tails
    (4, 32, 33, 40),            # D=32 fast path
    (2, 256, 40, 30),           # D=256 fast path
    (3, 16, 21, 25),            # generic-D path
    (2, 100, 19, 25),           # generic-D path, D not a multiple of 32
    (1, 128, 1, 2),             # K=1
    (3, 128, 300, 3),           # heavy index collisions (n=3 rows)
])
def test_kernels_vs_oracle_closed_form(pkg, co, B, D, K, n):
    from multimodal_learning_b200 import crd
    m1, m2, v1, v2, y, idx = _random_problem(B, D, K, n, seed=B * 1000 + D)
    T = float(torch.tensor(0.07).item())
    d = lambda t: t.to(DEV)
    # first-call Z, exactly like the reference: mean(raw)*n
    params = torch.tensor([K, T, -1, -1, 0.5], device=DEV)
    crd.crd_scores(d(m1), d(m2), d(v1), d(v2), d(idx), T, set_Z=params[2:4], want_out=False)
    raw1, raw2 = co.contrast_scores(m1, m2, torch.tensor([K, T, -1, -1, 0.5]), v1, v2, idx)
    Z1, Z2 = (raw1.mean() * n).item(), (raw2.mean() * n).item()
    assert abs(params[2].item() - Z1) / Z1 < TOL and abs(params[3].item() - Z2) / Z2 < TOL
    Zt = torch.tensor([Z1, Z2], device=DEV)
    loss, _, g1, g2, o1, o2 = crd.crd_fused_loss_grad(d(m1), d(m2), d(v1), d(v2), d(idx), T, Zt, n, K, want_out=True)
    w_loss, w_g1, w_g2, x1, x2 = co.crd_closed_form(m1, m2, v1, v2, idx, T, Z1, Z2, n)
    assert rel_err(loss, w_loss.reshape(1)) < TOL
    # gradients are sums of +/- terms of size ~ |row| / (T*B); when they cancel (e.g. n=2 and the negative IS the
    # positive's row) the exact result is ~0 and only an absolute comparison at that scale is meaningful
    gscale = 1.0 / (T * B) * max(m1.abs().max().item(), m2.abs().max().item())
    for got, want in ((g1, w_g1), (g2, w_g2)):
        err = (got.double().cpu() - want).abs().max().item()
        assert err < TOL * max(want.abs().max().item(), 1e-2 * gscale)
    assert rel_err(o1, x1) < TOL and rel_err(o2, x2) < TOL
    # scores-only kernel agrees with the fused kernel's optional outputs
    s1, s2, sums = crd.crd_scores(d(m1), d(m2), d(v1), d(v2), d(idx), T, Z=Zt, want_sums=True)
    assert torch.equal(s1, o1) and torch.equal(s2, o2)
    # weighted-rows kernel == einsum
    c1, c2 = torch.rand(B, K + 1), torch.rand(B, K + 1)
    h1, h2 = crd.crd_weighted_rows(d(m1), d(m2), d(idx), d(c1), d(c2))
    assert rel_err(h1, torch.einsum("bk,bkd->bd", c1.double(), m2.double()[idx])) < TOL
    assert rel_err(h2, torch.einsum("bk,bkd->bd", c2.double(), m1.double()[idx])) < TOL
    # row update vs oracle; untouched rows bit-identical
    b1, b2 = d(m1).clone(), d(m2).clone()
    crd.crd_memory_update(b1, b2, d(v1), d(v2), d(y), 0.5)
    r1, r2 = m1.clone(), m2.clone()
    co.momentum_update_(r1, y, v1, 0.5)
    co.momentum_update_(r2, y, v2, 0.5)
    assert rel_err(b1, r1) < 1e-6 and rel_err(b2, r2) < 1e-6
    untouched = torch.ones(n, dtype=torch.bool)
    untouched[y] = False
    assert torch.equal(b1.cpu()[untouched], m1[untouched]) and torch.equal(b2.cpu()[untouched], m2[untouched])


def test_ragged_segments_equal_dense_and_shard_sum(pkg, co):
    """Row-sharded bank in one process: each 'rank' owns a contiguous row block and scores only the
    (anchor, column) pairs whose row it owns (ragged segments, local row ids).  Summing the ranks'
    partial loss terms / gradients reproduces the dense single-bank result; owner-applies updates
    reproduce the dense update bit for bit."""
    from multimodal_learning_b200 import crd
    B, D, K, n, R = 6, 128, 200, 97, 3
    m1, m2, v1, v2, y, idx = _random_problem(B, D, K, n, seed=5)
    T = float(torch.tensor(0.07).item())
    d = lambda t: t.to(DEV)
    Zt = torch.tensor([31.0, 29.0], device=DEV)
    loss, _, g1, g2, o1, o2 = crd.crd_fused_loss_grad(d(m1), d(m2), d(v1), d(v2), d(idx), T, Zt, n, K, want_out=True)
    # dense expressed as ragged: identical bits
    seg = torch.arange(0, (B + 1) * (K + 1), K + 1, dtype=torch.int64)
    _, sums_r, g1r, g2r, _, _ = crd.crd_fused_loss_grad(d(m1), d(m2), d(v1), d(v2), d(idx.reshape(-1)), T, Zt, n, K,
                                                       seg_ptr=d(seg), want_sums=True)
    assert torch.equal(g1r, g1) and torch.equal(g2r, g2)
    assert abs(-(sums_r[0] + sums_r[1]).item() / B - loss.item()) < 1e-5 * abs(loss.item())
    # shard rows over R ranks
    rows_per = (n + R - 1) // R
    tot_sums = torch.zeros(4, dtype=torch.float64)
    tg1 = torch.zeros(B, D, dtype=torch.float64)
    tg2 = torch.zeros(B, D, dtype=torch.float64)
    u1, u2 = m1.clone(), m2.clone()
    for r in range(R):
        lo, hi = r * rows_per, min(n, (r + 1) * rows_per)
        own = (idx >= lo) & (idx < hi)
        counts = own.sum(1)
        segp = torch.zeros(B + 1, dtype=torch.int64)
        segp[1:] = counts.cumsum(0)
        local = (idx[own] - lo).contiguous()                    # row-major order keeps column 0 first
        posf = own[:, 0].to(torch.uint8)
        if local.numel() == 0:
            continue
        _, sums, pg1, pg2, _, _ = crd.crd_fused_loss_grad(
            d(m1[lo:hi].contiguous()), d(m2[lo:hi].contiguous()), d(v1), d(v2), d(local), T, Zt, n, K,
            seg_ptr=d(segp), pos_flag=d(posf), batch_norm=B, want_sums=True)
        tot_sums += sums.double().cpu()
        tg1 += pg1.double().cpu()
        tg2 += pg2.double().cpu()
        s1, s2 = d(m1[lo:hi].contiguous()), d(m2[lo:hi].contiguous())
        crd.crd_memory_update(s1, s2, d(v1), d(v2), d(y), 0.5, row_begin=lo, row_end=hi)
        u1[lo:hi], u2[lo:hi] = s1.cpu(), s2.cpu()
    assert abs(-(tot_sums[0] + tot_sums[1]).item() / B - loss.item()) < TOL * abs(loss.item())
    assert rel_err(tg1, g1) < TOL and rel_err(tg2, g2) < TOL
    f1, f2 = d(m1).clone(), d(m2).clone()
    crd.crd_memory_update(f1, f2, d(v1), d(v2), d(y), 0.5)
    assert torch.equal(f1.cpu(), u1) and torch.equal(f2.cpu(), u2)


@pytest.mark.parametrize("W,Bl,D,K,n,chunk,skew", [
    (8, 4, 128, 9000, 4000, 2048, False),    # config-5-like: 8 owners, CTAs walk several slots, staged in shared memory
    (2, 6, 128, 700, 301, 256, False),       # two owners, one slot per CTA
    (8, 1, 64, 9000, 64, 2048, True),        # every id owned by rank 0: 4 full slots overflow the staging buffer -> direct walk
    (3, 2, 48, 90, 50, 32, False),           # generic-D kernel in peer mode
])
def test_peer_kernels_emulated_world_on_one_gpu(pkg, co, W, Bl, D, K, n, chunk, skew):
    """The NVLink-pull kernels (`mml_shard_route_strided` -> `mml_crd_fused_loss_grad_peer`) with W 'ranks' emulated in
    one process: every rank's routed arena is an ordinary local buffer, each owner scores all W*Bl anchors against its
    row block, and the owners' partial sums add up to the dense single-bank result (integer routing exact: every
    (anchor, column) pair lands in exactly one owner's slot, positives first)."""
    import ctypes
    from multimodal_learning_b200 import _cabi, crd
    lib = _cabi.lib()
    Bg = W * Bl
    m1, m2, v1, v2, y, idx = _random_problem(Bg, D, K, n, seed=W * 100 + Bl)
    rows_per = (n + W - 1) // W
    if skew:
        idx = idx % rows_per
        y = idx[:, 0].clone()
    T = float(torch.tensor(0.07).item())
    d = lambda t: t.to(DEV)
    Zt = torch.tensor([31.0, 29.0], device=DEV)
    loss, _, g1, g2, _, _ = crd.crd_fused_loss_grad(d(m1), d(m2), d(v1), d(v2), d(idx), T, Zt, n, K)
    cols = K + 1
    chunks = (cols + chunk - 1) // chunk
    st = _cabi.cur_stream(torch.device(DEV))
    ids = [torch.full((W, Bl, chunks, chunk), -1, dtype=torch.int32, device=DEV) for _ in range(W)]     # per SOURCE rank
    cnt = [torch.zeros(W, Bl, chunks, dtype=torch.int32, device=DEV) for _ in range(W)]
    for s in range(W):
        ci = d(idx[s * Bl:(s + 1) * Bl].contiguous())
        _cabi.check(lib.mml_shard_route_strided(_cabi.dptr(ci), Bl, cols, chunk, rows_per, W, _cabi.dptr(cnt[s]),
                                                _cabi.dptr(ids[s]), st), "route")
    # routing is an exact stable partition
    for s in range(W):
        c = cnt[s].cpu()
        assert int(c.sum()) == Bl * cols
        for o in range(W):
            lo = o * rows_per
            for bl in range(Bl):
                want = idx[s * Bl + bl]
                got = torch.cat([ids[s][o, bl, ch, :c[o, bl, ch]].cpu() for ch in range(chunks)]).long() + lo
                assert torch.equal(got, want[(want >= lo) & (want < lo + rows_per)])
    V1, V2 = d(v1), d(v2)
    tot = torch.zeros(4, dtype=torch.float64)
    tg1 = torch.zeros(Bg, D, dtype=torch.float64)
    tg2 = torch.zeros(Bg, D, dtype=torch.float64)
    P = ctypes.c_void_p
    for o in range(W):
        lo, hi = o * rows_per, min(n, (o + 1) * rows_per)
        if hi <= lo:
            continue
        b1, b2 = d(m1[lo:hi].contiguous()), d(m2[lo:hi].contiguous())
        ids_ptrs = (P * W)(*[ids[s][o].data_ptr() for s in range(W)])
        cnt_ptrs = (P * W)(*[cnt[s][o].data_ptr() for s in range(W)])
        pos_flag = d(((idx[:, 0] >= lo) & (idx[:, 0] < hi)).to(torch.uint8))
        nws = lib.mml_crd_peer_workspace_bytes(Bg, chunks, D)
        ws = torch.empty(nws, dtype=torch.uint8, device=DEV)
        sums = torch.empty(4, device=DEV)
        pg1, pg2 = torch.empty(Bg, D, device=DEV), torch.empty(Bg, D, device=DEV)
        _cabi.check(lib.mml_crd_fused_loss_grad_peer(
            _cabi.dptr(b1), _cabi.dptr(b2), hi - lo, D, _cabi.dptr(V1), _cabi.dptr(V2), ids_ptrs, cnt_ptrs, W, Bl, chunks,
            chunk, _cabi.dptr(pos_flag), T, _cabi.dptr(Zt), n, K, Bg, _cabi.dptr(sums), _cabi.dptr(pg1), _cabi.dptr(pg2),
            _cabi.dptr(ws), nws, st), "fused_peer")
        tot += sums.double().cpu()
        tg1 += pg1.double().cpu()
        tg2 += pg2.double().cpu()
        # first-step statistics kernel over the same slots: raw exp sums
        s2 = torch.empty(4, device=DEV)
        _cabi.check(lib.mml_crd_scores_peer(_cabi.dptr(b1), _cabi.dptr(b2), hi - lo, D, _cabi.dptr(V1), _cabi.dptr(V2),
                                            ids_ptrs, cnt_ptrs, W, Bl, chunks, chunk, T, _cabi.dptr(s2), _cabi.dptr(ws), nws,
                                            st), "scores_peer")
        tot[2:] += s2.double().cpu()[2:]
    assert abs(-(tot[0] + tot[1]).item() / Bg - loss.item()) < TOL * abs(loss.item())
    assert rel_err(tg1, g1) < TOL and rel_err(tg2, g2) < TOL
    raw1, raw2 = co.contrast_scores(m1, m2, torch.tensor([K, T, -1, -1, 0.5]), v1, v2, idx)
    assert abs(tot[2].item() - raw1.double().sum().item()) < TOL * raw1.double().sum().item()
    assert abs(tot[3].item() - raw2.double().sum().item()) < TOL * raw2.double().sum().item()


def test_alias_draw_bit_exact_given_raw_draws(pkg, co, golden):
    """Same generator state => same indices: replay the two torch RNG calls the reference makes
    (:133 random_, :137 bernoulli) and push them through the oracle's select."""
    g = golden("alias")
    for c in g.cfg["cases"]:
        am = pkg.AliasMethod(torch.from_numpy(g.np(f"{c}.raw").copy()))
        am.cuda()
        N = 4096 + 17
        torch.manual_seed(123)
        out = am.draw(N)
        torch.manual_seed(123)
        kk = torch.zeros(N, dtype=torch.long, device=DEV).random_(0, am.alias.numel())
        b = torch.bernoulli(am.prob.index_select(0, kk))
        want = co.alias_select(kk.cpu().numpy(), b.cpu().numpy(), g.np(f"{c}.alias"))
        assert out.dtype == torch.int64 and np.array_equal(out.cpu().numpy(), want), c
        # golden raw draws (made on the reference's CPU generator) through the device select kernel
        kk_g, b_g = g.t(f"{c}.draw_kk", DEV), g.t(f"{c}.draw_b", DEV)
        sel = torch.empty_like(kk_g)
        from multimodal_learning_b200 import _cabi
        _cabi.check(_cabi.lib().mml_alias_select(_cabi.dptr(am.alias), _cabi.dptr(kk_g), _cabi.dptr(b_g), kk_g.numel(),
                                                 None, 1, _cabi.dptr(sel), _cabi.cur_stream(kk_g.device)), "select")
        assert np.array_equal(sel.cpu().numpy(), g.np(f"{c}.draw_out")), c


def test_sampled_idx_path_column0_is_anchor(pkg, golden):
    """idx=None branch (:37-39): indices are drawn on the device and column 0 is overwritten with y."""
    g = golden("crd_small")
    mod = _make_module(pkg, g.cfg, g.state_dict("init."))
    y = g.t("step0.idx", DEV)
    torch.manual_seed(9)
    loss = mod(g.t("step0.f_s", DEV), g.t("step0.f_t", DEV), y, None)
    assert torch.isfinite(loss).all()
    torch.manual_seed(9)
    drawn = mod.contrast.multinomial.draw(g.cfg["B"] * (g.cfg["K"] + 1), y=y, cols=g.cfg["K"] + 1).view(g.cfg["B"], -1)
    assert torch.equal(drawn[:, 0], y)
    assert int(drawn.min()) >= 0 and int(drawn.max()) < g.cfg["n"]


def test_full_size_config2_properties(pkg, co):
    """BASELINE config 2 (B=1024, D=128, n=1M, K=16384) -- too big for the CPU oracle in seconds, so:
    (i) a random subset of anchors is checked against the oracle at full K; (ii) the total loss equals the
    batch-normalised sum of two half-batch calls (linearity over anchors); (iii) updated rows are unit
    norm and every other row is bit-identical (checksum of the untouched part)."""
    from multimodal_learning_b200 import crd
    B, D, K, n = 1024, 128, 16384, 1_000_000
    gen = torch.Generator(device=DEV).manual_seed(1)
    stdv = 1.0 / (D / 3) ** 0.5
    m1 = torch.rand(n, D, device=DEV, generator=gen) * 2 * stdv - stdv
    m2 = torch.rand(n, D, device=DEV, generator=gen) * 2 * stdv - stdv
    v1 = torch.nn.functional.normalize(torch.randn(B, D, device=DEV, generator=gen), dim=1)
    v2 = torch.nn.functional.normalize(torch.randn(B, D, device=DEV, generator=gen), dim=1)
    y = torch.randperm(n, device=DEV, generator=gen)[:B]
    idx = torch.randint(0, n, (B, K + 1), device=DEV, generator=gen)
    idx[:, 0] = y
    T = float(torch.tensor(0.07).item())
    params = torch.tensor([K, T, -1, -1, 0.5], device=DEV)
    crd.crd_scores(m1, m2, v1, v2, idx, T, set_Z=params[2:4], want_out=False)
    Z = params[2:4].clone()
    loss, _, g1, g2, _, _ = crd.crd_fused_loss_grad(m1, m2, v1, v2, idx, T, Z, n, K)
    # (i) subset vs oracle closed form (CPU, float64)
    sub = torch.tensor([0, 1, 511, 1023], device=DEV)
    rows = idx[sub].cpu()
    uniq, inv = torch.unique(rows, return_inverse=True)
    _, w_g1, w_g2, _, _ = co.crd_closed_form(m1[uniq.to(DEV)].cpu(), m2[uniq.to(DEV)].cpu(), v1[sub].cpu(), v2[sub].cpu(),
                                            inv, T, Z[0].item(), Z[1].item(), n)
    # closed form divides by its own B (=4); the kernel divided by 1024
    assert rel_err(g1[sub] * (B / len(sub)), w_g1) < TOL
    assert rel_err(g2[sub] * (B / len(sub)), w_g2) < TOL
    # (i') loss, Z and both gradients of a 48-anchor slice at full K / n vs the float64 oracle: Z from the slice's own
    #      Monte-Carlo estimate (CRD_criterion.py:52-59), the loss normalised by the slice (batch_norm = 48)
    sub2 = torch.arange(200, 248, device=DEV)
    pz = torch.tensor([K, T, -1, -1, 0.5], device=DEV)
    v1s, v2s, idxs = v1[sub2].contiguous(), v2[sub2].contiguous(), idx[sub2].contiguous()
    crd.crd_scores(m1, m2, v1s, v2s, idxs, T, set_Z=pz[2:4], want_out=False)
    l_sub, _, gs1, gs2, _, _ = crd.crd_fused_loss_grad(m1, m2, v1s, v2s, idxs, T, Z, n, K)
    rows2 = idxs.cpu()
    uniq2, inv2 = torch.unique(rows2, return_inverse=True)
    w_l, w_s1, w_s2, x1, x2 = co.crd_closed_form(m1[uniq2.to(DEV)].cpu(), m2[uniq2.to(DEV)].cpu(), v1s.cpu(), v2s.cpu(), inv2, T,
                                                 Z[0].item(), Z[1].item(), n)
    assert abs(l_sub.item() - w_l.item()) < TOL * abs(w_l.item())
    assert rel_err(gs1, w_s1) < TOL and rel_err(gs2, w_s2) < TOL
    # x = exp(dot/T)/Z  =>  the slice's Z estimate is mean(x) * Z * n
    assert abs(pz[2].item() - x1.mean().item() * Z[0].item() * n) < TOL * pz[2].item()
    assert abs(pz[3].item() - x2.mean().item() * Z[1].item() * n) < TOL * pz[3].item()
    # ... and the full-batch loss is the batch-weighted sum of slice losses: (ii) below ties every slice to the full call
    l_chk, _, _, _, _, _ = crd.crd_fused_loss_grad(m1, m2, v1s, v2s, idxs, T, Z, n, K, batch_norm=B)
    assert abs(l_chk.item() * (B / len(sub2)) - l_sub.item()) < 1e-5 * abs(l_sub.item())
    # (ii) linearity over anchors
    h = B // 2
    la, _, ga1, _, _, _ = crd.crd_fused_loss_grad(m1, m2, v1[:h].contiguous(), v2[:h].contiguous(), idx[:h].contiguous(),
                                                  T, Z, n, K, batch_norm=B)
    lb, _, gb1, _, _, _ = crd.crd_fused_loss_grad(m1, m2, v1[h:].contiguous(), v2[h:].contiguous(), idx[h:].contiguous(),
                                                  T, Z, n, K, batch_norm=B)
    assert abs((la + lb).item() - loss.item()) < 1e-5 * abs(loss.item())
    # per-anchor work is independent; only the column chunking (a function of B) re-associates the sums
    assert rel_err(torch.cat([ga1, gb1]), g1) < 1e-5
    # (iii) update
    chk1 = m1.double().sum(1)
    crd.crd_memory_update(m1, m2, v1, v2, y, 0.5)
    assert (m1[y].norm(dim=1) - 1).abs().max().item() < 1e-5 and (m2[y].norm(dim=1) - 1).abs().max().item() < 1e-5
    mask = torch.ones(n, dtype=torch.bool, device=DEV)
    mask[y] = False
    assert torch.equal(m1.double().sum(1)[mask], chk1[mask])


def test_duplicate_anchor_ids_update_rows_once(pkg, co):
    """y with repeated ids (replacement sampling, DistributedSampler padding): the reference gathers all old rows first and
    index_copy_s the candidates (CRD_criterion.py:66-79); on the host that is 'the last occurrence wins'.  The kernel must
    leave exactly that well-formed, unit-norm row -- never a torn mix of two warps' writes."""
    from multimodal_learning_b200 import crd
    n, D, B = 50, 128, 64
    gen = torch.Generator().manual_seed(3)
    m1, m2 = co.memory_init(n, D, gen)
    v1 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=1)
    v2 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=1)
    y = torch.randint(0, 12, (B,), generator=gen)                      # 64 draws from 12 ids: every id repeats
    want1, want2 = m1.clone(), m2.clone()
    co.momentum_update_(want1, y, v1, 0.5)
    co.momentum_update_(want2, y, v2, 0.5)
    for _ in range(5):                                                 # racy code would differ from run to run
        d1, d2 = m1.to(DEV), m2.to(DEV)
        crd.crd_memory_update(d1, d2, v1.to(DEV), v2.to(DEV), y.to(DEV), 0.5)
        assert rel_err(d1, want1) < 1e-6 and rel_err(d2, want2) < 1e-6
        touched = torch.unique(y)
        assert (d1[touched.to(DEV)].norm(dim=1) - 1).abs().max().item() < 1e-5


def test_out_of_range_ids_are_flagged_not_dereferenced(pkg, golden):
    """The reference's index_select device-asserts on an id outside the bank (CRD_criterion.py:41,46).  The kernels clamp
    the id to row 0 -- no out-of-bounds read -- and raise a sticky flag that `check_device_errors()` turns into an IndexError."""
    g = golden("crd_d128")
    mod = _make_module(pkg, g.cfg, g.state_dict("init."))
    pkg.device_error_flags(reset=True)
    f_s, f_t = g.t("step0.f_s", DEV), g.t("step0.f_t", DEV)
    idx, cidx = g.t("step0.idx", DEV), g.t("step0.contrast_idx", DEV).clone()
    loss = mod(f_s, f_t, idx, cidx)
    assert torch.isfinite(loss).all()
    pkg.check_device_errors()                                   # clean run: nothing raised
    cidx[3, 5] = g.cfg["n"] + 12345                              # beyond the bank
    cidx[7, 9] = -4                                              # negative
    loss = mod(f_s, f_t, idx, cidx)
    assert torch.isfinite(loss).all()
    assert pkg.device_error_flags(reset=False) & pkg._cabi.DEVERR_CRD_INDEX
    with pytest.raises(IndexError):
        pkg.check_device_errors()
    assert pkg.device_error_flags() == 0                         # the check cleared the flag


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_modules_work_on_a_device_that_is_not_current(pkg, golden):
    """The reference's modules run wherever their tensors live; the C ABI acts on the current device, so the ctypes layer
    switches to the tensors' device for the duration of a call (and back)."""
    g = golden("crd_d128")
    c = g.cfg
    dev1 = torch.device("cuda", 1)
    assert torch.cuda.current_device() == 0
    mod = _make_module(pkg, c, g.state_dict("init.")).to(dev1)
    p = "step0."
    f_s = g.t(p + "f_s", dev1).requires_grad_(True)
    loss = mod(f_s, g.t(p + "f_t", dev1), g.t(p + "idx", dev1), g.t(p + "contrast_idx", dev1))
    loss.backward()
    assert torch.cuda.current_device() == 0
    assert rel_err(loss.reshape(-1), g.t(p + "loss")) < 1e-4 and rel_err(f_s.grad, g.t(p + "grad_f_s")) < 1e-4
