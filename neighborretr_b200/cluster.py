"""Global-feature producer: two levels of density-peak token merging per modality, the step right before the
head's ``global_level`` / centrality / Sinkhorn.

Reference: ``NeighborRetr.merge_global_features`` (NeighborRetr/models/modeling.py:446-481) wiring ``CTM`` and
``TCBlock`` of NeighborRetr/models/cluster.py (:453-561 clustering + merging, :638-717 CTM, :780-965 attention
block), instantiated at modeling.py:186-197 with ratios 1/6, 1/4 (text) and 1/4, 1/3 (video), k = 3, 8 heads.

This is a host-level PyTorch module, written from scratch on explicit tensors (the reference threads mutable
token dictionaries through the layers): it is trainable, encoder-side and tiny (O(B·N²·D)), so it stays library
ops — there is no kernel of ours here and it is not part of the measured hot path.  What it must keep is
(i) the parameter names, so that reference checkpoints load (``text_ctm0.conv.conv.weight``, ``text_ctm0.norm.*``,
``text_ctm0.score.*``, ``text_block0.norm1.*``, ``text_block0.attn.{q,kv,proj}.*``, same for ``*1`` and
``video_*``), and (ii) the arithmetic, including two things that are easy to miss in the reference:
* the CTM's score tensor is overwritten IN PLACE with -inf at masked tokens (``masked_fill_`` on a view,
  cluster.py:701-703), and that same tensor is the additive attention bias of the following block (:856, :881-883);
* the "far away" distance given to masked tokens uses the maximum over the WHOLE batch (:476-477), and the
  density tie-break noise is drawn with ``torch.rand`` in the order text-level0, video-level0, text-level1,
  video-level1 (cluster.py:483-484 via modeling.py:470-477).  ``noise=`` injects it for deterministic tests.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


def density_peak_clusters(x, cluster_num, k, mask=None, noise=None):
    """DPC-KNN assignment (reference cluster.py:453-509), no gradient.  x [B,N,C], mask [B,N] (>0 = real token)
    or None, noise [B,N] in [0,1) or None (then drawn with torch.rand).  Returns idx_cluster int64 [B,N]."""
    with torch.no_grad():
        B, N, C = x.shape
        dist = torch.cdist(x, x) / (C ** 0.5)
        keep = None
        if mask is not None:
            keep = mask > 0
            dist = dist * keep[:, None, :] + (dist.max() + 1) * (~keep[:, None, :])
        near = torch.topk(dist, k=k, dim=-1, largest=False).values
        density = (-(near ** 2).mean(dim=-1)).exp()
        if noise is None:
            noise = torch.rand(density.shape, device=density.device, dtype=density.dtype)
        density = density + noise * 1e-6
        if keep is not None:
            density = density * keep
        # distance to the nearest token of higher density (the sample's largest distance if there is none)
        higher = (density[:, None, :] > density[:, :, None]).type(x.dtype)
        far = dist.flatten(1).max(dim=-1).values[:, None, None]
        parent_dist = (dist * higher + far * (1 - higher)).min(dim=-1).values
        centres = torch.topk(parent_dist * density, k=cluster_num, dim=-1).indices            # [B,K]
        rows = torch.gather(dist, 1, centres[:, :, None].expand(B, cluster_num, N))           # dist[b, centre_j, :]
        idx_cluster = rows.argmin(dim=1)
        # a centre always belongs to its own cluster
        idx_cluster.scatter_(1, centres, torch.arange(cluster_num, device=x.device)[None, :].expand(B, cluster_num))
    return idx_cluster


def merge_by_cluster(x, idx_cluster, cluster_num, weight):
    """Weighted mean of the tokens of every cluster (reference cluster.py:512-546).  x [B,N,C], weight [B,N,1]
    (>= 0), idx_cluster [B,N] -> [B,K,C]; differentiable in x and weight."""
    B, N, C = x.shape
    flat = (idx_cluster + torch.arange(B, device=x.device)[:, None] * cluster_num).reshape(B * N)
    total = weight.new_zeros(B * cluster_num, 1)
    total.index_add_(0, flat, weight.reshape(B * N, 1))
    total = total + 1e-6
    share = weight / total[flat].reshape(B, N, 1)
    merged = x.new_zeros(B * cluster_num, C)
    merged.index_add_(0, flat, (x * share).reshape(B * N, C).type(x.dtype))
    return merged.reshape(B, cluster_num, C)


class TokenConv(nn.Module):
    """Residual 1-D convolution along the token axis (reference cluster.py:638-667)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, bias=False, padding=0):
        super().__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, bias=bias, padding=padding)

    def forward(self, x):
        return x + self.conv(x.transpose(1, 2)).transpose(1, 2)


class CTM(nn.Module):
    """Clustering token merger (reference cluster.py:670-717): project + score the tokens, cluster them by density
    peaks, merge every cluster into one token weighted by exp(score)."""

    def __init__(self, sample_ratio, embed_dim, dim_out, k=5):
        super().__init__()
        self.sample_ratio = sample_ratio
        self.dim_out = dim_out
        self.conv = TokenConv(embed_dim, dim_out, kernel_size=3, bias=False, padding=1)
        self.norm = nn.LayerNorm(dim_out)
        self.score = nn.Linear(dim_out, 1)
        self.k = k

    def forward(self, x, mask=None, noise=None):
        """x [B,N,C], mask [B,N] {0,1} or None -> (merged [B,K,C], tokens [B,N,C], score [B,N,1]).  ``score`` is
        -inf at masked tokens, as the reference's in-place fill leaves it for the attention bias."""
        tokens = self.norm(self.conv(x))
        score = self.score(tokens)
        if mask is not None:
            score = score.masked_fill((1 - mask).to(torch.bool).unsqueeze(2), float("-inf"))
        weight = score.exp()
        cluster_num = max(math.ceil(tokens.shape[1] * self.sample_ratio), 1)
        idx_cluster = density_peak_clusters(tokens, cluster_num, self.k, mask, noise)
        return merge_by_cluster(tokens, idx_cluster, cluster_num, weight), tokens, score


class TCAttention(nn.Module):
    """Multi-head attention of merged tokens over the un-merged ones with the token scores as additive bias
    (reference cluster.py:780-888, sr_ratio = 1)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None):
        super().__init__()
        if dim % num_heads:
            raise ValueError(f"dim {dim} should be divided by num_heads {num_heads}.")
        self.dim, self.num_heads = dim, num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, q_in, kv_in, bias):
        B, Nq, C = q_in.shape
        H, hd = self.num_heads, C // self.num_heads
        q = self.q(q_in).reshape(B, Nq, H, hd).permute(0, 2, 1, 3).contiguous()
        kv = self.kv(kv_in).reshape(B, -1, 2, H, hd).permute(2, 0, 3, 1, 4).contiguous()
        k, v = kv[0], kv[1]
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = (attn + bias.squeeze(-1)[:, None, None, :]).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, Nq, C))


class TCBlock(nn.Module):
    """Pre-norm attention block without MLP (reference cluster.py:891-965): merged + attn(norm(merged), norm(tokens))."""

    def __init__(self, dim, num_heads, qkv_bias=True, qk_scale=None):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = TCAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale)
        self.apply(_init_block_weights)

    def forward(self, merged, tokens, score):
        return merged + self.attn(self.norm1(merged), self.norm1(tokens), score)


def _init_block_weights(m):
    """Reference cluster.py:922-935: truncated normal (std 0.02, cut at +-2) for Linear, unit LayerNorm."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02, a=-2., b=2.)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


# (sample ratio level 0, level 1) per modality — reference modeling.py:186-197
RATIOS = {"text": (1 / 6, 1 / 4), "video": (1 / 4, 1 / 3)}


def init_token_clustering(module, dim=512, num_heads=8, k=3):
    """Register the eight merging layers on ``module`` under the reference's attribute names."""
    for mod, (r0, r1) in RATIOS.items():
        setattr(module, f"{mod}_ctm0", CTM(sample_ratio=r0, embed_dim=dim, dim_out=dim, k=k))
        setattr(module, f"{mod}_block0", TCBlock(dim=dim, num_heads=num_heads))
        setattr(module, f"{mod}_ctm1", CTM(sample_ratio=r1, embed_dim=dim, dim_out=dim, k=k))
        setattr(module, f"{mod}_block1", TCBlock(dim=dim, num_heads=num_heads))


def merge_global_features(module, text_feat, video_feat, text_mask, video_mask, noise=None):
    """``NeighborRetr.merge_global_features`` (reference modeling.py:446-481): [B,Nt,D], [B,Nv,D] ->
    ([B,Gt,D], [B,Gv,D]); Gt = Gv = 1 for Nt <= 24 words and Nv <= 12 frames.  ``noise``: optional 4-tuple of
    [B,N] tensors (text level 0, video level 0, text level 1, video level 1) replacing the torch.rand draws."""
    n = noise if noise is not None else (None,) * 4
    # level 0 (text first, then video: the order of the reference's random draws), masks apply here only —
    # merged tokens carry no mask (cluster.py:553-559)
    t_merged, t_tokens, t_score = module.text_ctm0(text_feat, text_mask.detach(), n[0])
    v_merged, v_tokens, v_score = module.video_ctm0(video_feat, video_mask.detach(), n[1])
    t = module.text_block0(t_merged, t_tokens, t_score)
    v = module.video_block0(v_merged, v_tokens, v_score)
    # level 1
    t_merged, t_tokens, t_score = module.text_ctm1(t, None, n[2])
    v_merged, v_tokens, v_score = module.video_ctm1(v, None, n[3])
    return module.text_block1(t_merged, t_tokens, t_score), module.video_block1(v_merged, v_tokens, v_score)
