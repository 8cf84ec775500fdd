"""Shared helpers of the parity tests: build the CUDA head and the oracle on identical seeded inputs."""
import os

import numpy as np
import torch

from neighborretr_b200 import synth
from oracle import head as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOG100 = float(np.log(100.0))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def make_head(d, cfg, params, precision="fp32", bwd_precision=None, device="cuda", fused=True):
    from neighborretr_b200.modeling import NeighborRetr
    m = NeighborRetr(cfg, width=d)
    for name, sd in params.items():
        getattr(m, name).load_state_dict(sd)
    m.clip.logit_scale.data.fill_(LOG100)
    m.head_precision = precision
    m.head_bwd_precision = bwd_precision
    m.head_fused = fused
    return m.to(device)


def set_bank(m, bank, device="cuda"):
    m.mb_ind = bank.mb_ind.to(device)
    m.mb_feat_t = bank.mb_feat_t.to(device)
    m.mb_feat_v = bank.mb_feat_v.to(device)
    m.mb_mask_t = bank.mb_mask_t.to(device)
    m.mb_mask_v = bank.mb_mask_v.to(device)
    m.mb_batch = bank.mb_ind.shape[0]


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu().reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def oracle_losses(h, bank, params, cfg, dtype=torch.float32):
    """Oracle head fwd+bwd on CPU; returns (losses[5], grads dict, intermediates)."""
    cast = lambda t: t.to(dtype) if t.is_floating_point() else t
    text = cast(h.text_feat).clone().requires_grad_(True)
    video = cast(h.video_feat).clone().requires_grad_(True)
    gt = cast(h.global_text).clone().requires_grad_(True)
    gv = cast(h.global_video).clone().requires_grad_(True)
    p = {k: {n: cast(v).clone().requires_grad_(True) for n, v in sd.items()} for k, sd in params.items()}
    lsp = torch.tensor(LOG100, dtype=dtype, requires_grad=True)
    losses = O.compute_losses(text, video, h.text_mask, h.video_mask, cast(bank.mb_feat_t), cast(bank.mb_feat_v),
                              bank.mb_mask_t, bank.mb_mask_v, gt, gv, p, lsp.exp(), cfg)
    losses[0].backward()
    grads = dict(text=text.grad, video=video.grad, gt=gt.grad, gv=gv.grad, logit_scale=lsp.grad)
    for nme in ("text_weight_fc", "video_weight_fc"):
        for pn, t in p[nme].items():
            grads[f"{nme}.{pn}"] = t.grad
    return torch.stack([x.detach() for x in losses]), grads


def cuda_losses(m, h, cfg):
    """CUDA head fwd+bwd; returns (losses[5], grads dict)."""
    hd = h.to("cuda")
    text = hd.text_feat.clone().requires_grad_(True)
    video = hd.video_feat.clone().requires_grad_(True)
    gt = hd.global_text.clone().requires_grad_(True)
    gv = hd.global_video.clone().requires_grad_(True)
    m.zero_grad(set_to_none=True)
    losses = m._compute_losses(text, video, hd.text_mask, hd.video_mask, m.mb_feat_t, m.mb_feat_v, m.mb_mask_t,
                               m.mb_mask_v, cfg.centrality_scale, cfg.beta, cfg.num_neighbors, cfg.temperature,
                               m.clip.logit_scale.exp(), global_feats=(gt, gv))
    losses[0].backward()
    grads = dict(text=text.grad, video=video.grad, gt=gt.grad, gv=gv.grad, logit_scale=m.clip.logit_scale.grad)
    for nme in ("text_weight_fc", "video_weight_fc"):
        for pn, t in getattr(m, nme).named_parameters():
            grads[f"{nme}.{pn}"] = t.grad
    return torch.stack([x.detach() for x in losses]).cpu(), {k: v.detach().cpu() for k, v in grads.items()}
