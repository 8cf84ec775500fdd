"""The tcgen05 GEMM of the token-weight MLPs (csrc/gemm_tc.cu) against float64 products of the SAME bf16 operands:
forward (K-major x K-major, bias + ReLU + second-layer dot in the epilogue), dx (K-major x MN-major),
dW1 (MN-major x MN-major, split-K over the tokens).  Operands are consumed as stored — no transposed copies — so these
tests pin the MN-major shared-memory descriptors.  Tolerance: fp32 accumulation of bf16 products, 2e-5 relative to the
largest entry (h is stored as bf16: 2^-9 relative)."""
import ctypes

import pytest
import torch

from neighborretr_b200 import ops
from neighborretr_b200.ops import _call, _p, _stream

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).contiguous()


@pytest.mark.parametrize("T,D,H", [(300, 512, 1024), (15360, 512, 1024), (128, 256, 512), (1000, 64, 256)])
def test_mlp_forward_gemm(T, D, H):
    g = torch.Generator().manual_seed(T)
    x = _bf(torch.randn(T, D, generator=g).cuda()); w1 = _bf((0.05 * torch.randn(H, D, generator=g)).cuda())
    b1 = (0.1 * torch.randn(H, generator=g)).cuda(); w2 = (0.05 * torch.randn(H, generator=g)).cuda()
    h = torch.empty(T, H, dtype=torch.bfloat16, device="cuda")
    logits = torch.zeros(T, device="cuda")
    _call("nr_mlp_fwd", _p(x), T, D, _p(w1), H, _p(b1), _p(w2), _p(h), _p(logits), _stream())
    want = torch.relu(x.double() @ w1.double().t() + b1.double())
    herr = (h.double() - want).abs().max().item() / want.abs().max().item()
    lwant = want @ w2.double()
    lerr = (logits.double() - lwant).abs().max().item() / lwant.abs().max().item()
    print(f"mlp_fwd T={T} D={D} H={H}: h rel err {herr:.2e}, logits rel err {lerr:.2e}")
    assert herr < 5e-3 and lerr < 2e-5
    # evaluation form: no hidden activations kept
    logits2 = torch.zeros(T, device="cuda")
    _call("nr_mlp_fwd", _p(x), T, D, _p(w1), H, _p(b1), _p(w2), None, _p(logits2), _stream())
    assert (logits2.double() - lwant).abs().max().item() / lwant.abs().max().item() < 2e-5


@pytest.mark.parametrize("T,D,H", [(300, 512, 1024), (3072, 512, 1024), (64, 256, 512)])
@pytest.mark.parametrize("acc", [0, 1])
def test_mlp_backward_dx_gemm(T, D, H, acc):
    g = torch.Generator().manual_seed(T + acc)
    dh = _bf(torch.randn(T, H, generator=g).cuda()); w1 = _bf((0.05 * torch.randn(H, D, generator=g)).cuda())
    dx = torch.zeros(T, D, device="cuda") if acc else torch.full((T, D), float("nan"), device="cuda")
    _call("nr_mlp_bwd_dx", _p(dh), T, H, _p(w1), D, _p(dx), acc, _stream())
    want = dh.double() @ w1.double()
    err = (dx.double() - want).abs().max().item() / want.abs().max().item()
    print(f"mlp_bwd_dx T={T} acc={acc}: rel err {err:.2e}")
    assert err < 2e-5


@pytest.mark.parametrize("T,D,H", [(300, 512, 1024), (15360, 512, 1024), (7680, 512, 1024), (200, 256, 512)])
def test_mlp_backward_dw1_gemm(T, D, H):
    g = torch.Generator().manual_seed(T)
    dh = _bf(torch.randn(T, H, generator=g).cuda()); x = _bf(torch.randn(T, D, generator=g).cuda())
    dw1 = torch.zeros(H, D, device="cuda")
    _call("nr_mlp_bwd_dw1", _p(dh), T, H, _p(x), D, _p(dw1), _stream())
    want = dh.double().t() @ x.double()
    err = (dw1.double() - want).abs().max().item() / want.abs().max().item()
    print(f"mlp_bwd_dw1 T={T}: rel err {err:.2e}")
    assert err < 2e-5


def test_token_softmax_and_cast():
    g = torch.Generator().manual_seed(1)
    ra, rb, n = 7, 5, 24
    logits = torch.randn(ra + rb, n, generator=g).cuda(); b2 = torch.tensor([0.3]).cuda()
    ma = (torch.rand(ra, n, generator=g) > 0.4).long().cuda(); mb = (torch.rand(rb, n, generator=g) > 0.4).long().cuda()
    ma[:, 0] = 1; mb[:, 0] = 1
    w = torch.empty(ra + rb, n, device="cuda")
    _call("nr_token_softmax", _p(logits), _p(b2), _p(ma), _p(mb), ra, ra + rb, n, _p(w), _stream())
    m = torch.cat([ma, mb])
    want = torch.softmax((logits.double() + 0.3).masked_fill(m == 0, -9e15), -1)
    assert (w.double() - want).abs().max().item() < 1e-6 and w[m == 0].abs().max().item() == 0.0
    x = torch.randn(1000 * 512 + 3, generator=g).cuda()[: 1000 * 512]
    y = torch.empty(x.numel(), dtype=torch.bfloat16, device="cuda")
    _call("nr_cast_bf16", _p(x.contiguous()), _p(y), x.numel(), _stream())
    assert torch.equal(y, x.to(torch.bfloat16))


@pytest.mark.parametrize("M,K,N,trans", [(128, 128, 512, 0), (128, 128, 512, 1), (40, 40, 64, 1), (1024, 1024, 512, 0),
                                         (70, 33, 20, 0), (33, 70, 20, 1)])
def test_small_fp32_matmul(M, K, N, trans):
    g = torch.Generator().manual_seed(M + K)
    A = torch.randn((K, M) if trans else (M, K), generator=g).cuda(); X = torch.randn(K, N, generator=g).cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    _call("nr_matmul_f32", _p(A), A.shape[1], trans, _p(X), N, M, K, N, _p(out), N, 0, _stream())
    want = (A.double().t() if trans else A.double()) @ X.double()
    assert (out.double() - want).abs().max().item() < 1e-5 * want.abs().max().item()
    _call("nr_matmul_f32", _p(A), A.shape[1], trans, _p(X), N, M, K, N, _p(out), N, 1, _stream())
    assert (out.double() - 2 * want).abs().max().item() < 2e-5 * want.abs().max().item()


def test_small_matvec():
    g = torch.Generator().manual_seed(0)
    A = torch.randn(5, 4, generator=g).cuda(); x = torch.randn(8, generator=g).cuda(); y = torch.randn(5, generator=g).cuda()
    o = torch.empty(5, device="cuda"); ot = torch.empty(4, device="cuda")
    _call("nr_matvec_small", _p(A), 5, 4, 0, _p(x), _p(x[4:]), _p(o), _stream())
    _call("nr_matvec_small", _p(A), 5, 4, 1, _p(y), None, _p(ot), _stream())
    assert torch.allclose(o, A @ (x[:4] + x[4:]), atol=1e-6) and torch.allclose(ot, A.t() @ y, atol=1e-6)


def test_token_weights_pair_matches_the_two_single_nodes():
    """Both modalities in one node (one cast launch, one grouped forward GEMM, one grouped backward GEMM) against the
    per-modality nodes on the library GEMMs, forward and every gradient (same bf16 arithmetic: 2e-2 on the gradients
    that sum thousands of cancelling terms, 1e-3 elsewhere)."""
    d, nt, nv, ra, rb = 256, 24, 12, 11, 17
    g = torch.Generator().manual_seed(8)
    mk = lambda *sh: torch.randn(*sh, generator=g).cuda()
    xt, xtb, xv, xvb = mk(ra, nt, d), mk(rb, nt, d), mk(ra, nv, d), mk(rb, nv, d)
    masks = [(torch.rand(r, n, generator=g) > 0.3).long().cuda() for r, n in ((ra, nt), (rb, nt), (ra, nv), (rb, nv))]
    for m in masks:
        m[:, 0] = 1
    ps = [[(0.05 * torch.randn(2 * d, d, generator=g)).cuda(), (0.05 * torch.randn(2 * d, generator=g)).cuda(),
           (0.05 * torch.randn(1, 2 * d, generator=g)).cuda(), (0.05 * torch.randn(1, generator=g)).cuda()] for _ in range(2)]
    ups = [torch.randn(r, n, generator=g).cuda() for r, n in ((ra, nt), (rb, nt), (ra, nv), (rb, nv))]

    def run(pair):
        leaves = [xt.clone().requires_grad_(True), xv.clone().requires_grad_(True)] + \
                 [t.clone().requires_grad_(True) for p_ in ps for t in p_]
        pt, pv = tuple(leaves[2:6]), tuple(leaves[6:10])
        if pair:
            outs = ops.token_weights_pair(pt, pv, leaves[0], masks[0], leaves[1], masks[2], "bf16", xtb, masks[1], xvb, masks[3])
        else:
            ops.USE_OWN_GEMM = False
            try:
                outs = (*ops.token_weights(pt, leaves[0], masks[0], "bf16", xtb, masks[1]),
                        *ops.token_weights(pv, leaves[1], masks[2], "bf16", xvb, masks[3]))
            finally:
                ops.USE_OWN_GEMM = True
        sum((o * u).sum() for o, u in zip(outs, ups)).backward()
        return [o.detach() for o in outs], [l.grad for l in leaves]

    o1, g1 = run(True)
    o2, g2 = run(False)
    for a, b in zip(o1, o2):
        assert (a - b).abs().max().item() < 2e-3
    names = ["dxt", "dxv", "w1t", "b1t", "w2t", "b2t", "w1v", "b1v", "w2v", "b2v"]
    errs = {}
    for n, a, b in zip(names, g1, g2):
        if n.startswith("b2"):
            assert a.abs().max().item() < 1e-4
            continue
        errs[n] = ((a.double() - b.double()).norm() / b.double().norm()).item()
    print("token_weights_pair vs single nodes:", errs)
    assert all(v < 3e-2 for v in errs.values()), errs
