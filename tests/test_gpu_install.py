"""The drop-in route a trainer takes: a from-scratch stand-in with the reference model's shape (its own encoder behind
``get_text_video_feat``, ``merge_global_features``, the token-weight MLPs, ``clip.logit_scale``, mb_* attributes;
reference NeighborRetr/models/modeling.py:137-197,251-312), the head rebound onto it by neighborretr_b200.bind_head
(what install() does to the reference class), then the reference trainer's call order (training/trainer.py:84-119):

    loss, *parts = model(text_ids, text_mask, video, video_mask, idx, step, None); loss.backward(); optimizer.step()

Checked against the oracle (CPU restatement of the reference head) step by step, bank FIFO included, for the
CUDA-graph replay path (head_graph=True: one replay per step, gradients handed to autograd) and the eager path."""
import numpy as np
import pytest
import torch
from torch import nn

from helpers import LOG100, rel_l2
from neighborretr_b200 import bind_head, synth
from oracle import head as O

pytestmark = pytest.mark.gpu
D, NT, NV, B, M = 512, 24, 12, 32, 64


class _Clip(nn.Module):
    def __init__(self):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.tensor(LOG100))


class StandIn(nn.Module):
    """Same attribute surface as the reference model; the 'encoders' are two small trainable maps so that the head's
    feature gradients are observable as encoder-parameter gradients."""

    def __init__(self, config, params):
        super().__init__()
        self.config = config
        self.clip = _Clip()
        for name in ("text_weight_fc", "video_weight_fc"):
            mlp = nn.Sequential(nn.Linear(D, 2 * D), nn.ReLU(inplace=True), nn.Linear(2 * D, 1))
            mlp.load_state_dict(params[name])
            setattr(self, name, mlp)
        self.text_scale = nn.Parameter(torch.ones(D))
        self.video_scale = nn.Parameter(torch.ones(D))
        self.global_gain = nn.Parameter(torch.tensor(1.0))
        self.mb_ind = torch.tensor([], dtype=torch.long)
        self.mb_feat_t = torch.empty((0, 0, 0)); self.mb_feat_v = torch.empty((0, 0, 0))
        self.mb_mask_t = torch.empty((0, 0)); self.mb_mask_v = torch.empty((0, 0))
        self.mb_batch = 0
        self.globals_in = None

    def get_text_video_feat(self, text_ids, text_mask, video, video_mask, shaped=False):
        # text_ids / video carry precomputed embeddings in this stand-in (the CLIP towers are out of scope)
        assert video.dim() == 4                                  # flattened [b * frames, C, H, W] like the reference
        b = text_ids.shape[0]
        tf, vf = text_ids.view(b, NT, D) * self.text_scale, video.view(b, NV, D) * self.video_scale
        if tf.requires_grad:
            tf.retain_grad(); vf.retain_grad()
        self.feats = (tf, vf)                                    # the test reads the head's feature gradients here
        return tf, vf

    def merge_global_features(self, text_feat, video_feat, text_mask, video_mask):
        gt, gv = self.globals_in
        return gt * self.global_gain, gv * self.global_gain


def _oracle_run(steps, bank, params, cfg, lr):
    """The same training loop on the oracle: plain SGD on every parameter, FIFO after each step."""
    # (*_fc1 only feed softmaxes over a single global token, which are identically 1: no gradient, kept constant)
    p = {k: {n: v.clone().requires_grad_(k in ("text_weight_fc", "video_weight_fc")) for n, v in sd.items()}
         for k, sd in params.items()}
    ts = torch.ones(D, requires_grad=True); vs = torch.ones(D, requires_grad=True)
    gg = torch.tensor(1.0, requires_grad=True); lsp = torch.tensor(LOG100, requires_grad=True)
    mb = dict(ind=bank.mb_ind.clone(), t=bank.mb_feat_t.clone(), v=bank.mb_feat_v.clone(), mt=bank.mb_mask_t.clone(),
              mv=bank.mb_mask_v.clone())
    leaves = [ts, vs, gg, lsp] + [t for sd in p.values() for t in sd.values() if t.requires_grad]
    out = []
    for h in steps:
        text, video = h.text_feat * ts, h.video_feat * vs
        text.retain_grad(); video.retain_grad()
        losses = O.compute_losses(text, video, h.text_mask, h.video_mask, mb["t"], mb["v"], mb["mt"], mb["mv"],
                                  h.global_text * gg, h.global_video * gg, p, lsp.exp(), cfg)
        for t in leaves:
            t.grad = None
        losses[0].backward()
        out.append((torch.stack([x.detach() for x in losses]), ts.grad.clone(), vs.grad.clone(), gg.grad.clone(),
                    p["text_weight_fc"]["0.weight"].grad.clone(), lsp.grad.clone(), text.grad.clone(), video.grad.clone()))
        with torch.no_grad():
            for t in leaves:
                t -= lr * t.grad
            mb["ind"] = torch.cat([h.idx, mb["ind"]])[:M]
            mb["t"] = torch.cat([text.detach(), mb["t"]])[:M]; mb["v"] = torch.cat([video.detach(), mb["v"]])[:M]
            mb["mt"] = torch.cat([h.text_mask, mb["mt"]])[:M]; mb["mv"] = torch.cat([h.video_mask, mb["mv"]])[:M]
    return out, mb


@pytest.mark.parametrize("graph,precision", [(True, "bf16x3"), (False, "bf16x3"), (True, "bf16"), (True, "fp32")])
def test_trainer_loop_through_the_rebound_forward(graph, precision):
    cfg = synth.default_config()
    params = synth.make_mlp_params(d=D)
    bank = synth.make_bank(M, NT, NV, d=D)
    steps = [synth.make_batch(B, NT, NV, d=D, seed=500 + i) for i in range(3)]
    lr = 1e-3
    want, mb_want = _oracle_run(steps, bank, params, cfg, lr)

    class Model(StandIn):
        pass
    names = bind_head(Model, graph=graph, precision=precision)
    assert "forward" in names and "head_forward" in names and "_compute_losses" in names
    model = Model(cfg, params).cuda().train()
    for n in ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v"):      # as MemoryBankManager does: assignment
        setattr(model, n, getattr(bank, n).cuda())
    model.mb_batch = M
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    # feature gradients: fp32 exact to rounding; bf16x3 within the arg-max near-tie floor of an fp32-class head (one
    # flipped route of ~10^5 moves rel-L2 by several 1e-3 at B = 32); bf16 as in test_zz_fullsize.  The encoder-side
    # [D] gradients (text_scale, video_scale) sum ~800 token gradients that largely cancel, so they are only compared
    # where the arithmetic is fp32-accurate.
    ltol, gtol = {"fp32": (1e-4, 2e-3), "bf16x3": (1e-4, 1.5e-2), "bf16": (1e-2, 8e-2)}[precision]
    for i, h in enumerate(steps):
        hd = h.to("cuda")
        model.globals_in = (hd.global_text, hd.global_video)
        video = hd.video_feat.view(B, NV, 1, 16, 32)                     # [b, frames, C, H, W]
        loss, c_, u_, n_, k_ = model(hd.text_feat.view(B, NT * D), hd.text_mask, video, hd.video_mask, hd.idx, i, None)
        assert loss.requires_grad and (not graph or not c_.requires_grad)
        loss.backward()
        got = torch.stack([loss.detach(), c_.detach(), u_.detach(), n_.detach(), k_.detach()]).cpu()
        w_l, w_ts, w_vs, w_gg, w_w1, w_ls, w_dt, w_dv = want[i]
        np.testing.assert_allclose(got.numpy(), w_l.numpy(), rtol=ltol, err_msg=f"step {i}")
        errs = {"text_feat": rel_l2(model.feats[0].grad, w_dt), "video_feat": rel_l2(model.feats[1].grad, w_dv),
                "global_gain": rel_l2(model.global_gain.grad, w_gg),
                "w1": rel_l2(model.text_weight_fc[0].weight.grad, w_w1),
                "logit_scale": rel_l2(model.clip.logit_scale.grad, w_ls)}
        if precision != "bf16":
            errs.update({"text_scale": rel_l2(model.text_scale.grad, w_ts), "video_scale": rel_l2(model.video_scale.grad, w_vs)})
        print(f"step {i} graph={graph} {precision}: losses {got.tolist()} grad rel-L2 {errs}")
        if precision == "bf16":
            # B = 32 with k = 20 neighbours: the 3e-4 similarity error of bf16 operands moves entries across the top-k
            # boundary of the neighbour loss, a discrete change of the gradient (measured 27 % on step 1 with BOTH
            # the fused and the one-direction kernels, tools/debug_bankmask.py) while the losses stay within 1e-3:
            # only the losses and the smooth gradients are asserted in this mode
            errs = {k: v for k, v in errs.items() if k in ("global_gain", "logit_scale")}
        for k, e in errs.items():
            assert e < (gtol if k != "w1" else max(gtol, 4e-2)), (i, k, e)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1e9)
        opt.step()
        opt.zero_grad()
    # the bank after three steps: FIFO of the (updated-encoder) features, newest first (reference modeling.py:235-249)
    assert torch.equal(model.mb_ind.cpu(), mb_want["ind"])
    assert torch.equal(model.mb_mask_t.cpu(), mb_want["mt"])
    tol = 1e-5 if precision != "bf16" else 1e-2         # features come from parameters updated by this run's gradients
    np.testing.assert_allclose(model.mb_feat_t.cpu().numpy(), mb_want["t"].numpy(), rtol=0, atol=tol * 30)
    # evaluation mode returns None before the head (reference :271-273)
    model.eval()
    with torch.no_grad():
        assert model(hd.text_feat.view(B, NT * D), hd.text_mask, video, hd.video_mask, hd.idx) is None


def test_assigning_a_new_bank_between_graph_steps_is_picked_up():
    """MemoryBankManager re-assigns model.mb_* every epoch (reference utils/memory_bank.py:206-211): the captured step
    must see the new rows (identity check + copy into the static storage), not the bank it was captured with."""
    cfg = synth.default_config()
    params = synth.make_mlp_params(d=D)

    class Model(StandIn):
        pass
    bind_head(Model, graph=True, precision="bf16")
    model = Model(cfg, params).cuda().train()
    h = synth.make_batch(B, NT, NV, d=D, seed=9).to("cuda")
    model.globals_in = (h.global_text, h.global_video)
    args = (h.text_feat.view(B, NT * D), h.text_mask, h.video_feat.view(B, NV, 1, 16, 32), h.video_mask, h.idx)
    out = []
    for seed in (999, 4242, 999):
        bank = synth.make_bank(M, NT, NV, d=D, seed=seed)
        for n in ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v"):
            setattr(model, n, getattr(bank, n).cuda())
        loss = model(*args)[0]
        loss.backward()
        out.append(float(loss))
        model.zero_grad()
    # same bank -> same loss up to the step's float atomics (the weight MLP's second layer accumulates its dot products
    # with red.add: two replays may differ in the last bits; seen once in ~10 full-suite runs), another bank -> another loss
    same, other = abs(out[0] - out[2]), abs(out[0] - out[1])
    assert same <= 4e-6 * abs(out[0]) and other > 100 * same + 1e-6, out
