"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by ``neighborretr_b200`` (the product path).

CPU restatement, in plain functional PyTorch, of the reference NeighborRetr retrieval head
(zzezze/NeighborRetr).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module, and only as the checker / the
CPU baseline — never as the thing shipped.

Parity pin: the reference has no tests or golden vectors of its own, so this
restatement is pinned against outputs of the reference itself, executed in the build container:
``oracle/gen_golden.py`` imports ``/root/reference`` (read-only), runs the reference's functions
on seeded inputs and commits the results under ``tests/golden/``;
``tests/test_oracle_vs_golden.py`` checks every function below against them.

Each function cites the reference file:line it follows (paths relative to /root/reference).
All functions are dtype-generic (run them in float64 for "truth", float32 for "reference").
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

NEG = -9e15  # the reference's masking constant (modeling.py:486, until_module.py:77-81)


# --------------------------------------------------------------------------------------------
# a3: token-weight MLP + masked softmax      (modeling.py:139-153 definition, :485-492 use)
# --------------------------------------------------------------------------------------------
def token_weight_logits(feat, p):
    """Linear(D,2D)-ReLU-Linear(2D,1) on un-normalised token features -> [R, N]."""
    h = F.relu(F.linear(feat, p["0.weight"].to(feat.dtype), p["0.bias"].to(feat.dtype)))
    return F.linear(h, p["2.weight"].to(feat.dtype), p["2.bias"].to(feat.dtype)).squeeze(-1)


def token_weights(feat, mask, p):
    """softmax over tokens of the MLP logits with masked tokens filled with -9e15
    (modeling.py:485-487 / :490-492)."""
    logit = token_weight_logits(feat, p)
    logit = logit.masked_fill((1 - mask).to(torch.bool), NEG)
    return torch.softmax(logit, dim=-1)


# --------------------------------------------------------------------------------------------
# a1: local_level — masked max-sim late interaction      (modeling.py:483-514)
# --------------------------------------------------------------------------------------------
def maxsim_from_weights(text_feat, video_feat, text_mask, video_mask, tw, vw, normalize=True):
    """S[a,b] = 1/2 (sum_t tw[a,t] max_v R + sum_v vw[b,v] max_t R),
    R[a,b,t,v] = <t^_at, v^_bv> * tm[a,t] * vm[b,v]  (masked pairs are exactly 0 and take part
    in the max; modeling.py:495-512)."""
    if normalize:
        t = F.normalize(text_feat, dim=-1)
        v = F.normalize(video_feat, dim=-1)
    else:
        t, v = text_feat, video_feat
    r = torch.einsum("atd,bvd->abtv", t, v)
    if text_mask is not None:
        r = r * text_mask.to(r.dtype)[:, None, :, None]
    if video_mask is not None:
        r = r * video_mask.to(r.dtype)[None, :, None, :]
    t2v = torch.einsum("abt,at->ab", r.max(dim=-1)[0], tw)
    v2t = torch.einsum("abv,bv->ab", r.max(dim=-2)[0], vw)
    return (t2v + v2t) / 2.0


def local_level(text_feat, video_feat, text_mask, video_mask, params):
    """(S, S^T) exactly as modeling.py:483-514; ``params`` holds ``text_weight_fc`` /
    ``video_weight_fc`` state dicts."""
    tw = token_weights(text_feat, text_mask, params["text_weight_fc"])
    vw = token_weights(video_feat, video_mask, params["video_weight_fc"])
    s = maxsim_from_weights(text_feat, video_feat, text_mask, video_mask, tw, vw)
    return s, s.T


# --------------------------------------------------------------------------------------------
# a4: global_level      (modeling.py:516-539) — no normalisation, no masks, MLPs *_fc1
# --------------------------------------------------------------------------------------------
def global_level(gt, gv, params):
    tw = torch.softmax(token_weight_logits(gt, params["text_weight_fc1"]), dim=-1)
    vw = torch.softmax(token_weight_logits(gv, params["video_weight_fc1"]), dim=-1)
    g = maxsim_from_weights(gt, gv, None, None, tw, vw, normalize=False)
    return g, g.T


# --------------------------------------------------------------------------------------------
# a5: compute_centrality_weights      (modeling.py:403-430)
# --------------------------------------------------------------------------------------------
def centrality_weights(text_feat, video_feat, gt, gv, centrality_scale):
    d = text_feat.size(2)
    t = F.normalize(text_feat.reshape(-1, d), dim=-1)   # pads included, no mask (:419-420)
    v = F.normalize(video_feat.reshape(-1, d), dim=-1)
    gtn = F.normalize(gt.squeeze(1), dim=-1)
    gvn = F.normalize(gv.squeeze(1), dim=-1)
    ct = (gtn @ t.T).mean(dim=-1)
    cv = (gvn @ v.T).mean(dim=-1)
    return torch.exp(ct * centrality_scale), torch.exp(cv * centrality_scale)


# --------------------------------------------------------------------------------------------
# a6: CentralityWeightingLoss      (until_module.py:303-328)
# --------------------------------------------------------------------------------------------
def centrality_weighting_loss(x, w):
    return (-(torch.diag(F.log_softmax(x, dim=-1)) * w)).mean()


# --------------------------------------------------------------------------------------------
# a7: NeighborAdjustingLoss      (until_module.py:56-211)
# --------------------------------------------------------------------------------------------
def neighbor_topk(x, k):
    """Indices [B,k] of the k largest off-diagonal entries per row (until_module.py:108-119).
    The reference uses an unstable full sort; ties are broken here towards the LOWER column
    index (stable descending sort), which is the contract of the CUDA kernel too."""
    b = x.size(0)
    eye = torch.eye(b, dtype=torch.bool, device=x.device)
    xs = torch.where(eye, torch.full_like(x, NEG), x)
    idx = torch.sort(xs, dim=-1, descending=True, stable=True)[1]
    return idx[:, :k]


def _minmax_norm(x, ext):
    """Min-max normalisation with min/max taken over the NON-ext entries
    (until_module.py:65-86: ``mask == 0`` keeps the similarity)."""
    big = torch.full_like(x, 9e15)
    lo = torch.where(~ext, x, big).min(dim=-1, keepdim=True)[0]
    hi = torch.where(~ext, x, -big).max(dim=-1, keepdim=True)[0]
    return (x - lo) / (hi - lo)


def neighbor_adjusting_loss(x, xmb, k, tau):
    b = x.size(0)
    top = neighbor_topk(x, k)
    nbr = torch.zeros(b, b, dtype=torch.bool, device=x.device)
    nbr.scatter_(1, top, True)
    ext = nbr | torch.eye(b, dtype=torch.bool, device=x.device)
    c = xmb.sum(dim=-1) / xmb.size(-1)                       # :181
    cexp = c.unsqueeze(0).repeat(b, 1)                       # :182 (indexed by COLUMN)
    nx = _minmax_norm(x, ext)
    nc = _minmax_norm(cexp, ext)
    adj = torch.where(nbr, nx - nc, torch.full_like(x, NEG))  # :189-193
    pw = torch.softmax(adj * tau, dim=-1)                    # :147
    pw = torch.where(nbr, pw, torch.zeros_like(pw))
    pw = pw.clone()
    pw.fill_diagonal_(1.0)                                   # :157
    masked = torch.where(ext, x, torch.full_like(x, NEG))    # :199-203
    lp = F.log_softmax(masked, dim=-1) * pw
    row = -lp.sum(dim=-1) / pw.sum(dim=-1)
    return row.mean()


# --------------------------------------------------------------------------------------------
# a8: UniformRegularizationLoss      (until_module.py:214-291)
# --------------------------------------------------------------------------------------------
def sinkhorn_targets(g, beta, iters=50):
    with torch.no_grad():
        m, n = g.shape
        norm = torch.tensor(float(m + n), dtype=g.dtype, device=g.device).log().neg()
        u = torch.zeros(m, dtype=g.dtype, device=g.device)
        v = torch.zeros(n, dtype=g.dtype, device=g.device)
        for _ in range(iters):
            u = norm - torch.logsumexp(g + v.unsqueeze(0), dim=1)
            v = norm - torch.logsumexp(g + u.unsqueeze(1), dim=0)
        z = g + u.unsqueeze(1) + v.unsqueeze(0) - norm
    q = z.exp()
    eye = torch.zeros(g.shape, dtype=torch.float32, device=g.device)   # float32 identity (:260)
    eye.fill_diagonal_(1)
    return beta * q + (1 - beta) * eye


def uniform_regularization_loss(g, scale, beta=0.3, iters=50):
    t = sinkhorn_targets(g, beta, iters)
    return (-(F.log_softmax(g * scale, dim=-1) * t).sum(dim=-1)).mean()


# --------------------------------------------------------------------------------------------
# a9: KLDivergenceLoss      (until_module.py:339-359)
# --------------------------------------------------------------------------------------------
def kl_divergence_loss(g, s):
    return F.kl_div(F.log_softmax(g, dim=-1), F.softmax(s, dim=-1), reduction="mean")


# --------------------------------------------------------------------------------------------
# a14: _compute_losses with the (random) token merge bypassed: global features are inputs
#      (modeling.py:314-360)
# --------------------------------------------------------------------------------------------
def compute_losses(text_feat, video_feat, text_mask, video_mask,
                   mb_feat_t, mb_feat_v, mb_mask_t, mb_mask_v,
                   gt, gv, params, logit_scale, cfg):
    s, st = local_level(text_feat, video_feat, text_mask, video_mask, params)          # :319
    g, gtr = global_level(gt, gv, params)                                              # :437
    lu = (uniform_regularization_loss(g, cfg.temperature, cfg.beta)                    # :440-442
          + uniform_regularization_loss(gtr, cfg.temperature, cfg.beta)) / 2
    lkl = (kl_divergence_loss(g, s) + kl_divergence_loss(gtr, st)) / 2                 # :329-332
    wt, wv = centrality_weights(text_feat, video_feat, gt, gv, cfg.centrality_scale)   # :367
    lc = (centrality_weighting_loss(s * logit_scale, wt)                               # :372-380
          + centrality_weighting_loss(st * logit_scale, wv)) / 2
    mb_t2v = local_level(text_feat, mb_feat_v, text_mask, mb_mask_v, params)[0]        # :389
    mb_v2t = local_level(mb_feat_t, video_feat, mb_mask_t, video_mask, params)[1]      # :390
    ln = (neighbor_adjusting_loss(s, mb_v2t, cfg.num_neighbors, cfg.temperature)       # :393-401
          + neighbor_adjusting_loss(st, mb_t2v, cfg.num_neighbors, cfg.temperature)) / 2
    total = lc + lu * cfg.uniform_weight + ln * cfg.neighbor_weight + lkl * cfg.kl_weight
    return total, lc, lu, ln, lkl


# --------------------------------------------------------------------------------------------
# a13: memory bank FIFO      (modeling.py:222-249)
# --------------------------------------------------------------------------------------------
def update_memory_bank(bank, idx, text_feat, video_feat, text_mask, video_mask):
    """``bank`` = dict with mb_ind/mb_feat_t/mb_feat_v/mb_mask_t/mb_mask_v; returns the new dict."""
    if bank["mb_feat_v"].size(0) == 0:
        return dict(mb_ind=idx.clone(), mb_feat_t=text_feat.clone(), mb_feat_v=video_feat.clone(),
                    mb_mask_t=text_mask.clone(), mb_mask_v=video_mask.clone())
    cap = bank["mb_feat_v"].size(0)
    new = dict(mb_ind=torch.cat((idx, bank["mb_ind"])),
               mb_feat_t=torch.cat((text_feat, bank["mb_feat_t"])),
               mb_feat_v=torch.cat((video_feat, bank["mb_feat_v"])),
               mb_mask_t=torch.cat((text_mask, bank["mb_mask_t"])),
               mb_mask_v=torch.cat((video_mask, bank["mb_mask_v"])))
    return {k: v[:cap] for k, v in new.items()}


# --------------------------------------------------------------------------------------------
# a11: _run_on_single_gpu      (training/evaluator.py:21-63) — tiling does not change values
# --------------------------------------------------------------------------------------------
def eval_similarity(text_feat, video_feat, text_mask, video_mask, params, mini_batch=64):
    rows = []
    with torch.no_grad():
        for tf, tm in zip(torch.split(text_feat, mini_batch), torch.split(text_mask, mini_batch)):
            row = [local_level(tf, vf, tm, vm, params)[0].cpu().numpy()
                   for vf, vm in zip(torch.split(video_feat, mini_batch),
                                     torch.split(video_mask, mini_batch))]
            rows.append(np.concatenate(row, axis=-1))
    sim = np.concatenate(rows, axis=0)
    return sim, sim.T


# --------------------------------------------------------------------------------------------
# a10: AllGather semantics      (until_module.py:367-388) as a single-process statement
# --------------------------------------------------------------------------------------------
def allgather_reference(per_rank_tensors):
    """Forward = rank-ordered concatenation along dim 0; backward on rank r = rows
    [r*b, (r+1)*b) of the gathered gradient (no reduction)."""
    return torch.cat(list(per_rank_tensors), dim=0)
