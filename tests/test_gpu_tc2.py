"""Fused two-direction tcgen05 max-sim (nr_maxsim2_fwd) against a float64 torch emulation on the same bf16 operand
copies (masked tokens zeroed), i.e. against local_level's arithmetic (reference modeling.py:495-512) with only the
operand rounding shared.  Row-direction values are exact fp32 accumulations (tol 2e-5 abs like the one-direction
kernel); column-direction values additionally lose 3 mantissa bits of (v + 2) to the packed arg-max (<= 2e-6)."""
import pytest
import torch

from neighborretr_b200 import ops, synth

pytestmark = pytest.mark.gpu


def _emulate(X, Y, wx, wy):
    x = X.xn_bf16.double(); y = Y.xn_bf16.double()
    r = torch.einsum("axd,byd->abxy", x, y)
    px, ys = r.max(dim=3)                   # [a,b,x]
    py, xs = r.max(dim=2)                   # [a,b,y]
    h = torch.einsum("abx,ax->ab", px, wx.double()) + torch.einsum("aby,by->ab", py, wy.double())
    return h, px, ys, py, xs, r


def _weights(mask, seed):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(mask.shape, generator=g).cuda()
    logits = logits.masked_fill(mask == 0, -9e15)
    return torch.softmax(logits, -1)


def _check_problem(X, Y, wx, wy, alpha, out, out2, saved):
    h, px, ys, py, xs, r = _emulate(X, Y, wx, wy)
    p_x, y_s, p_y, x_s = saved
    assert (out.double() - alpha * h).abs().max().item() < 2e-5
    if out2 is not None:
        assert torch.equal(out2, out.t())
    assert (p_x.double() - px).abs().max().item() < 2e-5
    assert (p_y.double() - py).abs().max().item() < 2e-5
    # arg-max: identical except where the top two candidates are within accumulation noise / key truncation
    bad = y_s.long() != ys
    if bad.any():
        chosen = torch.gather(r, 3, y_s.long().unsqueeze(-1)).squeeze(-1)
        assert (px - chosen)[bad].abs().max().item() < 2e-6
    bad = x_s.long() != xs
    if bad.any():
        chosen = torch.gather(r, 2, x_s.long().unsqueeze(2)).squeeze(2)
        assert (py - chosen)[bad].abs().max().item() < 4e-6
    assert int(y_s.max()) < Y.n and int(x_s.max()) < X.n


@pytest.mark.parametrize("rx,nx,ry,ny", [(128, 24, 512, 12), (512, 12, 128, 24), (37, 24, 45, 12), (24, 64, 40, 64),
                                         (5, 8, 3, 16), (130, 32, 70, 48), (3, 4, 2, 4), (40, 12, 33, 12),
                                         (7, 16, 300, 8), (19, 48, 21, 32)])
def test_fused_forward_matches_emulation(rx, nx, ry, ny):
    d = 512
    hx = synth.make_batch(rx, nx, ny, d=d, seed=7).to("cuda")
    hy = synth.make_batch(ry, nx, ny, d=d, seed=8).to("cuda")
    X = ops.Prepared(hx.text_feat, bf16=True, mask=hx.text_mask)
    Y = ops.Prepared(hy.video_feat, bf16=True, mask=hy.video_mask)
    # masked tokens are zero rows of the operand copy, everything else is the RN rounding of xn
    ref = X.xn.to(torch.bfloat16) * hx.text_mask.unsqueeze(-1).to(torch.bfloat16)
    assert torch.equal(X.xn_bf16, ref)
    wx, wy = _weights(hx.text_mask, 1), _weights(hy.video_mask, 2)
    out = torch.full((rx, ry), float("nan"), device="cuda")
    out2 = torch.full((ry, rx), float("nan"), device="cuda")
    saved = ops.maxsim2_fwd([dict(X=X, Y=Y, wx=wx, wy=wy, alpha=0.5, out=out, strides=(ry, 1), out2=out2,
                                  strides2=(1, rx))])
    _check_problem(X, Y, wx, wy, 0.5, out, out2, saved[0])


def test_fused_forward_cta_pair_multicast(monkeypatch):
    """Opt-in CTA-pair variant (clusters of 2, TMA multicast of the shared Y box): same results."""
    monkeypatch.setenv("NR_TC2_PAIR", "1")
    test_fused_forward_matches_emulation(128, 24, 512, 12)
    test_fused_forward_matches_emulation(131, 24, 300, 12)       # odd number of X boxes: a dummy half pair-tile


def test_fused_forward_three_problems_one_launch():
    """The batch pair and the two bank pairs of a head step as one tile list."""
    d, nt, nv, b, m = 512, 24, 12, 50, 70
    hb = synth.make_batch(b, nt, nv, d=d, seed=11).to("cuda")
    hm = synth.make_batch(m, nt, nv, d=d, seed=12).to("cuda")
    T = ops.Prepared(hb.text_feat, bf16=True, mask=hb.text_mask)
    V = ops.Prepared(hb.video_feat, bf16=True, mask=hb.video_mask)
    MT = ops.Prepared(hm.text_feat, bf16=True, mask=hm.text_mask)
    MV = ops.Prepared(hm.video_feat, bf16=True, mask=hm.video_mask)
    tw, vw = _weights(hb.text_mask, 1), _weights(hb.video_mask, 2)
    tw_mb, vw_mb = _weights(hm.text_mask, 3), _weights(hm.video_mask, 4)
    S = torch.empty(b, b, device="cuda"); ST = torch.empty(b, b, device="cuda")
    A = torch.empty(b, m, device="cuda"); C = torch.empty(b, m, device="cuda")
    probs = [dict(X=T, Y=V, wx=tw, wy=vw, alpha=0.5, out=S, strides=(b, 1), out2=ST, strides2=(1, b)),
             dict(X=T, Y=MV, wx=tw, wy=vw_mb, alpha=0.5, out=A, strides=(m, 1)),
             dict(X=MT, Y=V, wx=tw_mb, wy=vw, alpha=0.5, out=C, strides=(1, m))]      # stored as [video, bank_t]
    saved = ops.maxsim2_fwd(probs)
    _check_problem(T, V, tw, vw, 0.5, S, ST, saved[0])
    _check_problem(T, MV, tw, vw_mb, 0.5, A, None, saved[1])
    _check_problem(MT, V, tw_mb, vw, 0.5, C.t(), None, saved[2])


def test_fused_forward_rejects_unsupported_shapes():
    hx = synth.make_batch(4, 6, 12, d=512, seed=1).to("cuda")
    X = ops.Prepared(hx.text_feat, bf16=True, mask=hx.text_mask)
    Y = ops.Prepared(hx.video_feat, bf16=True, mask=hx.video_mask)
    assert not ops.maxsim2_supported(6, 12, 512)
    out = torch.empty(4, 4, device="cuda")
    with pytest.raises(RuntimeError, match="unsupported shape"):
        ops.maxsim2_fwd([dict(X=X, Y=Y, wx=torch.ones(4, 6, device="cuda"), wy=torch.ones(4, 12, device="cuda"),
                              alpha=1.0, out=out, strides=(4, 1))])


@pytest.mark.parametrize("rx,nx,ry,ny,d", [(128, 24, 512, 12, 512), (512, 12, 128, 24, 512), (37, 24, 45, 12, 512),
                                           (24, 64, 40, 64, 512), (9, 8, 7, 16, 256), (130, 32, 70, 48, 64),
                                           (3, 4, 2, 4, 128)])
def test_fused_backward_matches_autograd_of_emulation(rx, nx, ry, ny, d):
    """nr_maxsim2_bwd (both sides) and nr_maxsim2_bwd_w against torch autograd through the float64 emulation on the
    same bf16 operands.  Differences: bf16 rounding of the routing coefficients and of the staged source tokens
    (rel-L2 < 8e-3; measured ~3e-3), weights exact to fp32 accumulation."""
    hx = synth.make_batch(rx, nx, ny, d=d, seed=17).to("cuda")
    hy = synth.make_batch(ry, nx, ny, d=d, seed=18).to("cuda")
    X = ops.Prepared(hx.text_feat, bf16=True, mask=hx.text_mask)
    Y = ops.Prepared(hy.video_feat, bf16=True, mask=hy.video_mask)
    wx, wy = _weights(hx.text_mask, 1), _weights(hy.video_mask, 2)
    out = torch.empty(rx, ry, device="cuda")
    p_x, y_s, p_y, x_s = ops.maxsim2_fwd([dict(X=X, Y=Y, wx=wx, wy=wy, alpha=0.5, out=out, strides=(ry, 1))])[0]
    g = torch.randn(rx, ry, generator=torch.Generator().manual_seed(5)).cuda()
    # autograd reference
    xd = X.xn_bf16.double().requires_grad_(True); yd = Y.xn_bf16.double().requires_grad_(True)
    wxd = wx.double().requires_grad_(True); wyd = wy.double().requires_grad_(True)
    r = torch.einsum("axd,byd->abxy", xd, yd)
    s = 0.5 * (torch.einsum("abx,ax->ab", r.max(dim=3).values, wxd) + torch.einsum("aby,by->ab", r.max(dim=2).values, wyd))
    (s * g.double()).sum().backward()
    for gt in (g, None):
        # second round: the same gradient passed transposed (dh strides swapped), as the v2t orientation does
        gg, sr, sc = (g, ry, 1) if gt is not None else (g.t().contiguous(), 1, rx)
        dx = torch.zeros_like(X.xn); dy = torch.zeros_like(Y.xn)
        ops.maxsim2_bwd(0, Y, wx, wy, y_s, x_s, gg, sr, sc, 0.5, rx, nx, ry, ny, d, dx)
        ops.maxsim2_bwd(1, X, wx, wy, y_s, x_s, gg, sr, sc, 0.5, rx, nx, ry, ny, d, dy)
        dwx = torch.zeros_like(wx); dwy = torch.zeros_like(wy)
        ops.maxsim2_bwd_w(p_x, p_y, gg, sr, sc, 0.5, rx, nx, ry, ny, dwx, dwy)
        mx = hx.text_mask.bool(); my = hy.video_mask.bool()
        for got, ref, m in ((dx, xd.grad, mx), (dy, yd.grad, my)):
            a, b = got.double()[m], ref[m]
            assert ((a - b).norm() / b.norm()).item() < 8e-3
        for got, ref in ((dwx, wxd.grad), (dwy, wyd.grad)):
            assert ((got.double() - ref).norm() / ref.norm()).item() < 1e-4


def test_maxsim_function_fused_vs_one_direction_kernels(monkeypatch):
    """ops.maxsim (the autograd op behind local_level) in bf16: the fused two-direction kernels against the
    one-direction kernels (two forward + four backward launches) on the same operand rounding; and the exact fp32
    path for the similarity itself.  (Gradients are not compared with fp32: on near-tied token pairs the bf16 and
    fp32 arg-max legitimately differ, which moves whole token vectors.)"""
    b, nt, nv, d = 48, 24, 12, 512
    h = synth.make_batch(b, nt, nv, d=d, seed=3).to("cuda")
    tw, vw = _weights(h.text_mask, 1), _weights(h.video_mask, 2)
    res = {}
    for name, prec, fused in (("fp32", "fp32", True), ("one", "bf16", False), ("fused", "bf16", True)):
        monkeypatch.setattr(ops, "USE_FUSED_MAXSIM", fused)
        t = h.text_feat.clone().requires_grad_(True); v = h.video_feat.clone().requires_grad_(True)
        a = tw.clone().requires_grad_(True); c = vw.clone().requires_grad_(True)
        S, ST = ops.maxsim(t, v, a, c, h.text_mask, h.video_mask, precision=prec)
        assert torch.equal(S.t(), ST)
        gS = torch.randn(b, b, generator=torch.Generator().manual_seed(9)).cuda()
        ((S * gS).sum() + (ST * gS).sum() * 0.5).backward()
        res[name] = (S.detach(), t.grad, v.grad, a.grad, c.grad)
    assert (res["fused"][0] - res["fp32"][0]).abs().max().item() < 2e-3
    assert (res["fused"][0] - res["one"][0]).abs().max().item() < 2e-5
    for i in range(1, 5):
        rel = ((res["fused"][i] - res["one"][i]).norm() / res["one"][i].norm()).item()
        assert rel < 1e-2, (i, rel)
    # masked tokens receive exactly no gradient
    assert res["fused"][1][h.text_mask == 0].abs().max().item() == 0.0
    assert res["fused"][2][h.video_mask == 0].abs().max().item() == 0.0


def _torch_token_weights(x, mask, w1, b1, w2, b2):
    logit = (torch.relu(x @ w1.t() + b1) @ w2.t() + b2).squeeze(-1)
    logit = logit.masked_fill(mask == 0, -9e15)
    return torch.softmax(logit, dim=-1)


@pytest.mark.parametrize("mode,tol_w,tol_g", [("fp32", 2e-6, 2e-4), ("tf32", 1e-3, 6e-2), ("bf16", 1e-2, 1.5e-1)])
def test_token_weights_node_vs_torch(mode, tol_w, tol_g):
    """ops.token_weights (GEMM with ReLU epilogue + nr_token_weights_fwd/_bwd; batch and bank tokens in one node)
    against the reference's op chain (modeling.py:148-153, 485-492) in float64, forward and every gradient.
    fp32 is the parity mode (measured 3e-7).  In the reduced-precision GEMM modes the weights stay within 7e-5 (tf32)
    of float64, but gradients move by 2-3 % (tf32) on these unit-variance random inputs because hidden units whose
    pre-activation lies within the GEMM rounding of zero flip their ReLU gate."""
    d, n, ra, rb = 256, 12, 9, 14
    g = torch.Generator().manual_seed(3)
    xa = torch.randn(ra, n, d, generator=g).cuda(); xb = torch.randn(rb, n, d, generator=g).cuda()
    ma = (torch.rand(ra, n, generator=g) > 0.3).long().cuda(); mb = (torch.rand(rb, n, generator=g) > 0.3).long().cuda()
    ma[:, 0] = 1; mb[:, 0] = 1
    w1 = (0.05 * torch.randn(2 * d, d, generator=g)).cuda(); b1 = (0.05 * torch.randn(2 * d, generator=g)).cuda()
    w2 = (0.05 * torch.randn(1, 2 * d, generator=g)).cuda(); b2 = (0.05 * torch.randn(1, generator=g)).cuda()
    ga = torch.randn(ra, n, generator=g).cuda(); gb = torch.randn(rb, n, generator=g).cuda()
    leaves = [t.clone().requires_grad_(True) for t in (xa, w1, b1, w2, b2)]
    wa, wb = ops.token_weights(tuple(leaves[1:]), leaves[0], ma, mode, xb, mb)
    ((wa * ga).sum() + (wb * gb).sum()).backward()
    ref = [t.double().clone().requires_grad_(True) for t in (xa, w1, b1, w2, b2)]
    ra_ = _torch_token_weights(ref[0], ma, *ref[1:]); rb_ = _torch_token_weights(xb.double(), mb, *ref[1:])
    ((ra_ * ga.double()).sum() + (rb_ * gb.double()).sum()).backward()
    errs = {"wa": (wa.double() - ra_).abs().max().item(), "wb": (wb.double() - rb_).abs().max().item()}
    assert wa[ma == 0].abs().max().item() == 0.0                      # masked tokens: weight exactly 0
    for name, got, want in zip(("x", "w1", "b1", "w2", "b2"), leaves, ref):
        if got.numel() == 1:        # b2: softmax is shift invariant, the true gradient is exactly 0
            assert got.grad.abs().max().item() < 1e-5 and want.grad.abs().max().item() < 1e-12
            continue
        errs["d" + name] = ((got.grad.double() - want.grad).norm() / want.grad.norm()).item()
    print("token_weights", mode, errs)
    assert errs["wa"] < tol_w and errs["wb"] < tol_w, errs
    assert all(v < tol_g for k, v in errs.items() if k.startswith("d")), errs
    # single-input form (local_level / evaluation)
    w_only, none = ops.token_weights((w1, b1, w2, b2), xa, ma, mode)
    assert none is None and (w_only.double() - ra_).abs().max().item() < tol_w


@pytest.mark.parametrize("ra,rb,d", [(128, 128, 512), (37, 70, 96), (256, 256, 512), (1000, 300, 512), (260, 257, 96)])
def test_gram_f32_matches_float64(ra, rb, d):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(ra, d, generator=g).cuda(); b = torch.randn(rb, d, generator=g).cuda()
    out = torch.empty(ra, rb, device="cuda"); outT = torch.empty(rb, ra, device="cuda")
    ops._call("nr_gram_f32", ops._p(a), ops._p(b), ra, rb, d, ops._p(out), ops._p(outT), ops._stream())
    ref = a.double() @ b.double().t()
    assert (out.double() - ref).abs().max().item() < 2e-4 and torch.equal(outT, out.t())
