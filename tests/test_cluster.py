"""Global-feature producer (token clustering; reference modeling.py:446-481 + cluster.py) against
golden vectors produced by the reference's own classes (tests/golden/cluster.npz, oracle/gen_golden.py:run_cluster).
It is a host-level torch module, so parity is checked on CPU (bit-level: same ops, same torch.rand stream) and,
with the tie-break noise injected, on the GPU; plus the standalone head running end to end without global_feats."""
import numpy as np
import pytest
import torch
from torch import nn

from helpers import load_golden
from neighborretr_b200 import cluster as C, synth
from oracle.gen_golden import CLUSTER_CASES


def _build(name, gold, device="cpu"):
    b, nt, nv, d, heads = CLUSTER_CASES[name]
    shell = nn.Module()
    C.init_token_clustering(shell, dim=d, num_heads=heads, k=3)
    sd = {k[len(name) + 3:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith(f"{name}_p_")}
    assert set(sd) == set(shell.state_dict()), set(sd) ^ set(shell.state_dict())      # the reference's parameter names
    shell.load_state_dict(sd)
    h = synth.make_batch(b, nt, nv, d=d, seed=2024)
    text = (h.text_feat / 6).clone().to(device).requires_grad_(True)
    video = (h.video_feat / 6).clone().to(device).requires_grad_(True)
    return shell.to(device), h, text, video


def _noise(name):
    """The four torch.rand draws of the reference after manual_seed(123), in its order."""
    b, nt, nv, _, _ = CLUSTER_CASES[name]
    import math
    kt, kv = max(math.ceil(nt / 6), 1), max(math.ceil(nv / 4), 1)
    torch.manual_seed(123)
    return tuple(torch.rand(b, n) for n in (nt, nv, kt, kv))


def _check(name, gold, gt, gv, text, video, shell, rtol, atol):
    wt = torch.linspace(-1, 1, gt.numel()).view_as(gt).to(gt.device)
    wv = torch.linspace(1, -1, gv.numel()).view_as(gv).to(gv.device)
    ((gt * wt).sum() + (gv * wv).sum()).backward()
    np.testing.assert_allclose(gt.detach().cpu().numpy(), gold[f"{name}_gt"], rtol=rtol, atol=atol)
    np.testing.assert_allclose(gv.detach().cpu().numpy(), gold[f"{name}_gv"], rtol=rtol, atol=atol)
    np.testing.assert_allclose(text.grad.cpu().numpy(), gold[f"{name}_dtext"], rtol=rtol, atol=atol)
    np.testing.assert_allclose(video.grad.cpu().numpy(), gold[f"{name}_dvideo"], rtol=rtol, atol=atol)
    for k, p in shell.named_parameters():
        want = gold[f"{name}_g_{k}"]
        if want.size == 0:
            assert p.grad is None, k
        else:
            np.testing.assert_allclose(p.grad.cpu().numpy(), want, rtol=rtol, atol=10 * atol, err_msg=k)


@pytest.mark.parametrize("name", list(CLUSTER_CASES))
def test_merge_global_features_matches_reference(name):
    gold = load_golden("cluster")
    shell, h, text, video = _build(name, gold)
    torch.manual_seed(123)                                   # the reference drew its noise from this state
    gt, gv = C.merge_global_features(shell, text, video, h.text_mask, h.video_mask)
    assert gt.shape == gold[f"{name}_gt"].shape and gv.shape == gold[f"{name}_gv"].shape
    _check(name, gold, gt, gv, text, video, shell, rtol=1e-5, atol=1e-6)
    # injected noise == the torch.rand stream
    shell.zero_grad(set_to_none=True)
    t2, v2 = text.detach().clone().requires_grad_(True), video.detach().clone().requires_grad_(True)
    gt2, gv2 = C.merge_global_features(shell, t2, v2, h.text_mask, h.video_mask, noise=_noise(name))
    assert torch.equal(gt2, gt) and torch.equal(gv2, gv)


def test_density_peak_clusters_properties():
    """Every centre keeps its own cluster id, every cluster is non-empty, masked tokens never become centres, and
    the merged token of a cluster is the exp(score)-weighted mean of its members."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 12, 16, generator=g)
    mask = (torch.arange(12)[None, :] < torch.tensor([12, 7, 3, 9, 5])[:, None]).long()
    idx = C.density_peak_clusters(x, 3, 3, mask, noise=torch.rand(5, 12, generator=g))
    assert idx.shape == (5, 12) and idx.min() == 0 and idx.max() == 2
    for b in range(5):
        assert set(idx[b].tolist()) == {0, 1, 2}
    w = torch.rand(5, 12, 1, generator=g) * mask[:, :, None]
    merged = C.merge_by_cluster(x, idx, 3, w)
    for b in range(5):
        for c in range(3):
            sel = idx[b] == c
            want = (x[b, sel] * w[b, sel]).sum(0) / (w[b, sel].sum() + 1e-6)
            torch.testing.assert_close(merged[b, c], want, rtol=1e-5, atol=1e-6)


def test_standalone_head_registers_reference_parameter_names():
    from neighborretr_b200.modeling import NeighborRetr
    m = NeighborRetr(synth.default_config(), width=32, token_clustering=True, cluster_heads=4)
    names = {n for n, _ in m.named_parameters()}
    for mod in ("text", "video"):
        for lvl in (0, 1):
            for suf in ("conv.conv.weight", "norm.weight", "norm.bias", "score.weight", "score.bias"):
                assert f"{mod}_ctm{lvl}.{suf}" in names
            for suf in ("norm1.weight", "norm1.bias", "attn.q.weight", "attn.q.bias", "attn.kv.weight", "attn.kv.bias",
                        "attn.proj.weight", "attn.proj.bias"):
                assert f"{mod}_block{lvl}.{suf}" in names
    h = synth.make_batch(8, 24, 12, d=32, seed=1)
    gt, gv = m.merge_global_features(h.text_feat, h.video_feat, h.text_mask, h.video_mask)
    assert gt.shape == (8, 1, 32) and gv.shape == (8, 1, 32)
    with pytest.raises(NotImplementedError):
        NeighborRetr(synth.default_config(), width=32).merge_global_features(h.text_feat, h.video_feat, h.text_mask,
                                                                            h.video_mask)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CLUSTER_CASES))
def test_gpu_merge_global_features_matches_reference(name):
    """Same module on the device (library ops) with the reference's noise injected: cluster assignments are decided
    by fp32 distances computed in a different order, so values are compared within 1e-4 instead of bitwise."""
    gold = load_golden("cluster")
    shell, h, text, video = _build(name, gold, device="cuda")
    hd = h.to("cuda")
    noise = tuple(n.cuda() for n in _noise(name))
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):      # cuDNN's conv1d defaults to TF32
        gt, gv = C.merge_global_features(shell, text, video, hd.text_mask, hd.video_mask, noise=noise)
        _check(name, gold, gt, gv, text, video, shell, rtol=1e-3, atol=1e-4)


@pytest.mark.gpu
def test_gpu_head_end_to_end_with_its_own_global_features():
    """The standalone head with token clustering enabled runs fwd+bwd with no global_feats passed in: the merged
    features feed the fused CUDA head, and their gradient reaches the clustering parameters."""
    from helpers import make_head, set_bank
    from neighborretr_b200.modeling import NeighborRetr
    b, nt, nv, d, mrows = 32, 24, 12, 64, 48
    cfg = synth.default_config()
    torch.manual_seed(0)
    m = NeighborRetr(cfg, width=d, token_clustering=True, cluster_heads=8).cuda()
    m.clip.logit_scale.data.fill_(float(np.log(100.0)))
    m.head_precision = "fp32"
    set_bank(m, synth.make_bank(mrows, nt, nv, d=d))
    h = synth.make_batch(b, nt, nv, d=d, seed=9).to("cuda")
    text = h.text_feat.clone().requires_grad_(True)
    noise = tuple(torch.rand(b, n, device="cuda") for n in (nt, nv, 4, 3))
    gfeats = m.merge_global_features(text, h.video_feat, h.text_mask, h.video_mask, noise=noise)
    assert gfeats[0].shape == (b, 1, d)
    losses = m._compute_losses(text, h.video_feat, h.text_mask, h.video_mask, m.mb_feat_t, m.mb_feat_v, m.mb_mask_t,
                               m.mb_mask_v, cfg.centrality_scale, cfg.beta, cfg.num_neighbors, cfg.temperature,
                               m.clip.logit_scale.exp(), global_feats=gfeats)
    assert all(torch.isfinite(x) for x in losses)
    losses[0].backward()
    assert text.grad is not None and torch.isfinite(text.grad).all()
    g = m.text_block1.attn.proj.weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0
    # same losses when the head calls merge_global_features itself (noise drawn by torch.rand from the same seed)
    torch.manual_seed(5)
    l1 = m._compute_losses(text.detach(), h.video_feat, h.text_mask, h.video_mask, m.mb_feat_t, m.mb_feat_v,
                           m.mb_mask_t, m.mb_mask_v, cfg.centrality_scale, cfg.beta, cfg.num_neighbors,
                           cfg.temperature, m.clip.logit_scale.exp())
    torch.manual_seed(5)
    l2 = m._compute_losses(text.detach(), h.video_feat, h.text_mask, h.video_mask, m.mb_feat_t, m.mb_feat_v,
                           m.mb_mask_t, m.mb_mask_v, cfg.centrality_scale, cfg.beta, cfg.num_neighbors,
                           cfg.temperature, m.clip.logit_scale.exp())
    torch.testing.assert_close(torch.stack(l1), torch.stack(l2), rtol=1e-5, atol=1e-6)
