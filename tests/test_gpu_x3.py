"""bf16x3: the fp32-accurate tensor-core mode (split-bf16 operands hi + lo, products hi.hi + lo.hi + hi.lo on the
tcgen05 kernels; forward with K = 3d, backward with three routing jobs per pair and exact fp32 routing coefficients).
Bars (north_star, fp32/tf32 class): losses 1e-4 relative, feature gradients rel-L2 5e-3, similarities abs 1e-5."""
import numpy as np
import pytest
import torch

from helpers import cuda_losses, load_golden, make_head, oracle_losses, rel_l2, set_bank
from neighborretr_b200 import ops, synth
from oracle.gen_golden import CASES, make_case

pytestmark = pytest.mark.gpu


def _emulate(text, video, tw, vw, tm, vm):
    """float64 local_level below the MLPs (reference modeling.py:495-512)."""
    t = torch.nn.functional.normalize(text.double(), dim=-1)
    v = torch.nn.functional.normalize(video.double(), dim=-1)
    r = torch.einsum("atd,bvd->abtv", t, v) * tm.double()[:, None, :, None] * vm.double()[None, :, None, :]
    t2v = (r.max(dim=3)[0] * tw.double()[:, None, :]).sum(2)
    v2t = (r.max(dim=2)[0] * vw.double()[None, :, :]).sum(2)
    return (t2v + v2t) / 2


@pytest.mark.parametrize("shape", [(40, 24, 12), (33, 12, 24), (17, 64, 64), (25, 8, 4)])
def test_maxsim_x3_forward_and_backward_vs_float64(shape):
    b, nt, nv = shape
    d = 512
    h = synth.make_batch(b, nt, nv, d=d, seed=91).to("cuda")
    g = torch.Generator().manual_seed(3)
    tw = torch.softmax(torch.randn(b, nt, generator=g), -1).cuda() * h.text_mask
    vw = torch.softmax(torch.randn(b, nv, generator=g), -1).cuda() * h.video_mask
    up = torch.randn(b, b, generator=g).cuda()
    res = {}
    for prec in ("bf16x3", "bf16"):
        text = h.text_feat.clone().requires_grad_(True); video = h.video_feat.clone().requires_grad_(True)
        twp = tw.clone().requires_grad_(True); vwp = vw.clone().requires_grad_(True)
        s, st = ops.maxsim(text, video, twp, vwp, h.text_mask, h.video_mask, prec)
        assert torch.equal(st, s.t())
        (s * up).sum().backward()
        res[prec] = (s.detach(), text.grad, video.grad, twp.grad, vwp.grad)
    text = h.text_feat.clone().requires_grad_(True); video = h.video_feat.clone().requires_grad_(True)
    twp = tw.clone().requires_grad_(True); vwp = vw.clone().requires_grad_(True)
    want = _emulate(text, video, twp, vwp, h.text_mask, h.video_mask)
    (want * up.double()).sum().backward()
    s3 = res["bf16x3"]
    serr = float((s3[0].double() - want).abs().max())
    gerr = [rel_l2(a, b_) for a, b_ in zip(s3[1:], (text.grad, video.grad, twp.grad, vwp.grad))]
    gerr_b = [rel_l2(a, b_) for a, b_ in zip(res["bf16"][1:], (text.grad, video.grad, twp.grad, vwp.grad))]
    print(f"x3 {shape}: S abs err {serr:.2e}; grad rel-L2 x3 {gerr} (bf16 {gerr_b})")
    assert serr < 1e-5
    assert max(gerr[:2]) < 2e-3 and max(gerr[2:]) < 1e-4


X3_CASES = {"cfg1": CASES["cfg1"], "tiny": dict(b=40, nt=8, nv=4, d=64, m=56, k=20)}


@pytest.mark.parametrize("name", ["tiny", "cfg1"])
def test_head_x3_vs_reference_golden_and_oracle(name):
    c = X3_CASES[name]
    gold = load_golden(name) if name in CASES else None
    h, bank, params, cfg = make_case(c)
    ref, rgrads = oracle_losses(h, bank, params, cfg)
    m = make_head(c["d"], cfg, params, "bf16x3")
    set_bank(m, bank)
    losses, grads = cuda_losses(m, h, cfg)
    lerr = float((losses / ref - 1).abs().max())
    gerr = {k: rel_l2(grads[k], rgrads[k]) for k in rgrads}
    print(f"head x3 [{name}]: loss rel err {lerr:.2e}; grad rel-L2 {gerr}")
    if gold is not None:
        np.testing.assert_allclose(losses.numpy(), gold["losses"], rtol=1e-4)      # the reference's own values
    np.testing.assert_allclose(losses.numpy(), ref.numpy(), rtol=1e-4)
    for k in ("text", "video", "gt", "gv"):
        assert gerr[k] < 5e-3, (k, gerr[k])
    assert gerr["logit_scale"] < 1e-3
    # token-weight MLP parameters: TF32 library GEMMs in this mode
    for k, e in gerr.items():
        if k.startswith(("text_weight_fc", "video_weight_fc")):
            assert e < 3e-2, (k, e)
