"""tcgen05 tensor-core max-sim (NR_PREC_BF16) against (i) a torch emulation that rounds the normalised tokens to
bf16 exactly like the operand copy does (isolates kernel errors from rounding: tol 2e-5 abs) and (ii) the fp32
CUDA-core kernel (rounding only: S abs err < 2e-3; the survey measured 1.9e-4)."""
import numpy as np
import pytest
import torch

from neighborretr_b200 import ops, synth
from neighborretr_b200._lib import NR_PREC_BF16, NR_PREC_FP32

pytestmark = pytest.mark.gpu


def _emulate(xn_bf16, yn_bf16, wx, mx, my):
    x = xn_bf16.double(); y = yn_bf16.double()
    r = torch.einsum("axd,byd->abxy", x, y)
    r = r * mx.double()[:, None, :, None] * my.double()[None, :, None, :]
    p, ys = r.max(dim=-1)
    return torch.einsum("abx,ax->ab", p, wx.double()), p, ys, r


def _dir(prec, X, Y, wx, mx, my):
    out = torch.empty(X.r, Y.r, dtype=torch.float32, device="cuda")
    p, ys = ops._maxsim_dir_fwd(prec, X, Y, wx, mx, my, 1.0, out, Y.r, 1, None, 0, 0, 0, True)
    return out, p, ys


@pytest.mark.parametrize("rx,nx,ry,ny", [(128, 24, 512, 12), (512, 12, 128, 24), (37, 24, 45, 12), (24, 64, 40, 64),
                                         (5, 8, 3, 16), (130, 32, 70, 48)])
def test_tc_forward_matches_bf16_emulation(rx, nx, ry, ny):
    d = 512
    hx = synth.make_batch(rx, nx, ny, d=d, seed=7).to("cuda")
    hy = synth.make_batch(ry, nx, ny, d=d, seed=8).to("cuda")
    X = ops.Prepared(hx.text_feat, bf16=True)
    Y = ops.Prepared(hy.video_feat, bf16=True)
    g = torch.Generator().manual_seed(1)
    wx = torch.softmax(torch.randn(rx, nx, generator=g), -1).cuda()
    mx, my = hx.text_mask, hy.video_mask
    assert torch.equal(X.xn_bf16, X.xn.to(torch.bfloat16))               # operand copy = RN rounding of xn
    h_tc, p_tc, y_tc = _dir(NR_PREC_BF16, X, Y, wx, mx, my)
    h_em, p_em, y_em, r = _emulate(X.xn_bf16, Y.xn_bf16, wx, mx, my)
    assert (p_tc.double() - p_em).abs().max().item() < 2e-5
    assert (h_tc.double() - h_em).abs().max().item() < 2e-5
    # arg-max byte: 255 ("no gradient") iff the X token is masked or the winning pair is a masked (zero) one;
    # otherwise identical to the emulation except where the top two candidates are within accumulation noise
    win_masked = torch.gather(my[None, :, None, :].expand(rx, ry, nx, ny), 3, y_em.unsqueeze(-1)).squeeze(-1) == 0
    nograd = win_masked | (mx[:, None, :] == 0)
    near_zero = p_em.abs() < 2e-6          # a real pair within noise of the masked zero may flip the flag
    assert torch.equal((y_tc == 255) | near_zero, nograd | near_zero)
    live = ~nograd & (y_tc != 255)
    bad = live & (y_tc.long() != y_em)
    if bad.any():
        chosen = torch.gather(r, 3, y_tc.long().clamp(max=ny - 1).unsqueeze(-1)).squeeze(-1)
        assert (p_em - chosen)[bad].abs().max().item() < 2e-6
    h_32, p_32, _ = _dir(NR_PREC_FP32, X, Y, wx, mx, my)
    assert (h_tc - h_32).abs().max().item() < 2e-3


def test_tc_forward_accumulate_and_transposed_output():
    d, rx, nx, ry, ny = 256, 20, 24, 33, 12
    hx = synth.make_batch(rx, nx, ny, d=d, seed=3).to("cuda")
    hy = synth.make_batch(ry, nx, ny, d=d, seed=4).to("cuda")
    X = ops.Prepared(hx.text_feat, bf16=True); Y = ops.Prepared(hy.video_feat, bf16=True)
    wx = torch.full((rx, nx), 1.0 / nx, device="cuda")
    base = torch.randn(rx, ry, device="cuda")
    out = base.clone(); out2 = base.t().contiguous()
    ops._maxsim_dir_fwd(NR_PREC_BF16, X, Y, wx, hx.text_mask, hy.video_mask, 0.5, out, ry, 1, out2, 1, rx, 1, False)
    ref = torch.empty(rx, ry, device="cuda")
    ops._maxsim_dir_fwd(NR_PREC_FP32, X, Y, wx, hx.text_mask, hy.video_mask, 0.5, ref, ry, 1, None, 0, 0, 0, False)
    assert (out - (base + ref)).abs().max().item() < 2e-3
    assert torch.equal(out2, out.t())


@pytest.mark.parametrize("rx,nx,ry,ny,d", [(128, 24, 512, 12, 512), (512, 12, 128, 24, 512), (37, 24, 45, 12, 512),
                                           (24, 64, 40, 64, 512), (9, 8, 7, 16, 256), (130, 32, 70, 48, 64)])
def test_tc_backward_matches_fp32_kernels(rx, nx, ry, ny, d):
    """tcgen05 backward contractions (generated routing tile x TMA-staged transposed source) vs the fp32
    gather/scatter kernels on the same arg-max: differences are bf16 rounding of coefficients and sources only
    (rel-L2 < 6e-3; measured ~3e-3)."""
    hx = synth.make_batch(rx, nx, ny, d=d, seed=17).to("cuda")
    hy = synth.make_batch(ry, nx, ny, d=d, seed=18).to("cuda")
    X = ops.Prepared(hx.text_feat, bf16=True)
    Y = ops.Prepared(hy.video_feat, bf16=True)
    g = torch.Generator().manual_seed(5)
    wx = torch.softmax(torch.randn(rx, nx, generator=g), -1).cuda()
    mx, my = hx.text_mask, hy.video_mask
    out = torch.empty(rx, ry, device="cuda")
    _, ys = ops._maxsim_dir_fwd(NR_PREC_FP32, X, Y, wx, mx, my, 1.0, out, ry, 1, None, 0, 0, 0, True)
    dH = torch.randn(rx, ry, generator=g).cuda()
    st = ops._stream()
    res = {}
    for prec in (NR_PREC_FP32, NR_PREC_BF16):
        dx = torch.zeros_like(X.xn); dy = torch.zeros_like(Y.xn)
        ysrc, yld = Y.bwd_source(prec); xsrc, xld = X.bwd_source(prec)
        ops._call("nr_maxsim_bwd_x", prec, ops._p(ysrc), yld, ops._p(wx), ops._p(mx), ops._p(my), ops._p(ys),
                  ops._p(dH), ry, 1, 0.7, rx, nx, ry, ny, d, ops._p(dx), st)
        ops._call("nr_maxsim_bwd_y", prec, ops._p(xsrc), xld, ops._p(wx), ops._p(mx), ops._p(my), ops._p(ys),
                  ops._p(dH), ry, 1, 0.7, rx, nx, ry, ny, d, ops._p(dy), st)
        res[prec] = (dx, dy)
    for a, b in zip(res[NR_PREC_BF16], res[NR_PREC_FP32]):
        rel = ((a - b).norm() / b.norm()).item()
        assert rel < 6e-3, rel
    # transposed dH strides (the v2t orientation) through the same kernels
    dHt = dH.t().contiguous()
    dx2 = torch.zeros_like(X.xn)
    ysrc, yld = Y.bwd_source(NR_PREC_BF16)
    ops._call("nr_maxsim_bwd_x", NR_PREC_BF16, ops._p(ysrc), yld, ops._p(wx), ops._p(mx), ops._p(my), ops._p(ys),
              ops._p(dHt), 1, rx, 0.7, rx, nx, ry, ny, d, ops._p(dx2), st)
    assert ((dx2 - res[NR_PREC_BF16][0]).norm() / res[NR_PREC_BF16][0].norm()).item() < 1e-5
