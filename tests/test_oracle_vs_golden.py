"""Pins the oracle (oracle/head.py, oracle/metrics.py) to outputs of the REFERENCE itself
(tests/golden/*.npz, produced by oracle/gen_golden.py from /root/reference).  CPU only."""
import os

import numpy as np
import pytest
import torch

from neighborretr_b200 import synth
from oracle import head as O
from oracle import metrics as OM
from oracle.gen_golden import CASES, make_case


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, f"{name}.npz")))


def _close(a, b, rtol=2e-5, atol=2e-6):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


def _sub(t, n=4096):
    f = t.reshape(-1)
    return f[:: max(1, f.numel() // n)]


@pytest.mark.parametrize("name", list(CASES))
def test_pieces_and_losses(golden_dir, name):
    c = CASES[name]
    gold = _load(golden_dir, name)
    h, bank, params, cfg = make_case(c)
    text = h.text_feat.clone().requires_grad_(True)
    video = h.video_feat.clone().requires_grad_(True)
    gt = h.global_text.clone().requires_grad_(True)
    gv = h.global_video.clone().requires_grad_(True)
    params = {k: {n: p.clone().requires_grad_(True) for n, p in sd.items()} for k, sd in params.items()}
    lsp = torch.tensor(float(np.log(100.0)), requires_grad=True)
    ls = lsp.exp()

    s, st = O.local_level(text, video, h.text_mask, h.video_mask, params)
    g, gT = O.global_level(gt, gv, params)
    wt, wv = O.centrality_weights(text, video, gt, gv, cfg.centrality_scale)
    mb_t2v = O.local_level(text, bank.mb_feat_v, h.text_mask, bank.mb_mask_v, params)[0]
    mb_v2t = O.local_level(bank.mb_feat_t, video, bank.mb_mask_t, h.video_mask, params)[1]
    _close(s.detach(), gold["S"]); _close(g.detach(), gold["G"], rtol=1e-5, atol=1e-5)
    _close(wt.detach(), gold["w_t"]); _close(wv.detach(), gold["w_v"])
    _close(mb_t2v.detach(), gold["mb_t2v"]); _close(mb_v2t.detach(), gold["mb_v2t"])

    _close(O.centrality_weighting_loss(s * ls, wt).item(), gold["Lc_t2v"], rtol=1e-5)
    _close(O.centrality_weighting_loss(st * ls, wv).item(), gold["Lc_v2t"], rtol=1e-5)
    _close(O.neighbor_adjusting_loss(s, mb_v2t, c["k"], cfg.temperature).item(), gold["Ln_t2v"], rtol=1e-5)
    _close(O.neighbor_adjusting_loss(st, mb_t2v, c["k"], cfg.temperature).item(), gold["Ln_v2t"], rtol=1e-5)
    _close(O.uniform_regularization_loss(g, cfg.temperature, cfg.beta).item(), gold["Lu_t2v"], rtol=1e-5)
    _close(O.uniform_regularization_loss(gT, cfg.temperature, cfg.beta).item(), gold["Lu_v2t"], rtol=1e-5)
    _close(O.kl_divergence_loss(g, s).item(), gold["Lkl_t2v"], rtol=1e-4, atol=1e-8)
    _close(O.kl_divergence_loss(gT, st).item(), gold["Lkl_v2t"], rtol=1e-4, atol=1e-8)
    _close(O.sinkhorn_targets(g.detach(), cfg.beta, 50), gold["sinkhorn_T"], rtol=1e-4, atol=1e-7)

    # top-k neighbour SET is bit-exact (no ties in these inputs)
    top = O.neighbor_topk(s.detach(), c["k"])
    nbr = torch.zeros_like(s, dtype=torch.uint8).scatter_(1, top, 1)
    assert np.array_equal(nbr.numpy(), gold["nbr_mask"])

    losses = O.compute_losses(text, video, h.text_mask, h.video_mask, bank.mb_feat_t, bank.mb_feat_v,
                              bank.mb_mask_t, bank.mb_mask_v, gt, gv, params, ls, cfg)
    _close(torch.stack([x.detach() for x in losses]), gold["losses"], rtol=1e-5)
    losses[0].backward()
    pick = _sub if int(gold["subsampled"]) else (lambda t: t)
    gtol = dict(rtol=2e-3, atol=2e-6)
    _close(pick(text.grad), gold["g_text"], **gtol)
    _close(pick(video.grad), gold["g_video"], **gtol)
    _close(pick(gt.grad), gold["g_gt"], **gtol)
    _close(pick(gv.grad), gold["g_gv"], **gtol)
    _close(lsp.grad, gold["g_logit_scale"], rtol=1e-4)
    _close(text.grad.norm(), gold["gn_text"], rtol=1e-4)
    _close(video.grad.norm(), gold["gn_video"], rtol=1e-4)
    for nme in ("text_weight_fc", "video_weight_fc"):
        for pn, p in params[nme].items():
            _close(p.grad.norm(), gold[f"gn_{nme}.{pn}"], rtol=1e-4, atol=1e-9)
            _close(pick(p.grad), gold[f"g_{nme}.{pn}"], rtol=2e-3, atol=1e-6)


def test_act_shaped_local_level(golden_dir):
    gold = _load(golden_dir, "act_piece")
    c = dict(b=24, nt=64, nv=64, d=512, m=8, k=20)
    h, bank, params, cfg = make_case(c)
    s, _ = O.local_level(h.text_feat, h.video_feat, h.text_mask, h.video_mask, params)
    _close(s, gold["S"])


def test_eval_similarity_and_metrics(golden_dir):
    gold = _load(golden_dir, "eval")
    c = dict(b=100, nt=8, nv=6, d=64, m=8, k=20)
    h, bank, params, cfg = make_case(c)
    sim, sim_t = O.eval_similarity(h.text_feat, h.video_feat[:70], h.text_mask, h.video_mask[:70], params)
    assert sim.dtype == np.float32 and sim.shape == (100, 70) and sim_t.shape == (70, 100)
    _close(sim, gold["sim_100x70"])
    for k in ("rand", "ties", "sim"):
        m = OM.compute_metrics(gold[f"{k}_mat"])
        assert np.array_equal(np.asarray(m["cols"]), gold[f"{k}_cols"])       # integer-exact
        np.testing.assert_array_equal(
            np.asarray([m["R1"], m["R5"], m["R10"], m["R50"], m["MR"], m["MedianR"], m["MeanR"]]),
            gold[f"{k}_scalars"])
        g_, e_ = OM.ranks_by_counting(gold[f"{k}_mat"])
        cols = np.concatenate([np.arange(a, a + b) for a, b in zip(g_, e_)])
        assert np.array_equal(cols, gold[f"{k}_cols"])                         # counting form == sort form


def test_memory_bank_fifo(golden_dir):
    gold = _load(golden_dir, "bank")
    c = dict(b=6, nt=4, nv=3, d=8)
    bank = dict(mb_ind=torch.tensor([], dtype=torch.long), mb_feat_t=torch.empty(0, 0, 0),
                mb_feat_v=torch.empty(0, 0, 0), mb_mask_t=torch.empty(0, 0), mb_mask_v=torch.empty(0, 0))
    for step in range(4):
        h = synth.make_batch(c["b"], c["nt"], c["nv"], d=c["d"], seed=50 + step, rank=step)
        if step == 1:
            bk = synth.make_bank(14, c["nt"], c["nv"], d=c["d"])
            bank = dict(mb_ind=bk.mb_ind, mb_feat_t=bk.mb_feat_t, mb_feat_v=bk.mb_feat_v,
                        mb_mask_t=bk.mb_mask_t, mb_mask_v=bk.mb_mask_v)
        bank = O.update_memory_bank(bank, h.idx, h.text_feat, h.video_feat, h.text_mask, h.video_mask)
        assert np.array_equal(bank["mb_ind"].numpy(), gold[f"ind_{step}"])
        assert np.array_equal(bank["mb_feat_v"].numpy(), gold[f"feat_v_{step}"])
        assert np.array_equal(bank["mb_mask_t"].numpy(), gold[f"mask_t_{step}"])
