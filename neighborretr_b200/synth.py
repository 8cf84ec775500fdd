"""Seeded synthetic inputs for the retrieval head.

Pair-correlated embeddings so that the losses are non-degenerate at logit_scale=100:
``base ~ N(0,I)``, ``text = base + sigma*N``, ``video = base + sigma*N`` with sigma=6, ragged
prefix-one int64 masks, a memory bank drawn the same way from seed 999 and global features
``3*normalize(base + 0.3*N)``.  Everything is drawn with a CPU ``torch.Generator`` so that the GPU
path, the oracle and the golden fixtures see bit-identical inputs on every machine.

This module is input plumbing only: no arithmetic of the head lives here.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace

import numpy as np
import torch

SHAPES = {
    # name: (Nt words, Nv frames, memory rows M)      BASELINE.json configs[0..2]
    "msrvtt": (24, 12, 512),
    "activitynet": (64, 64, 1024),
}


@dataclass
class HeadInputs:
    text_feat: torch.Tensor    # [b, Nt, D] fp32
    video_feat: torch.Tensor   # [b, Nv, D] fp32
    text_mask: torch.Tensor    # [b, Nt] int64 {0,1}
    video_mask: torch.Tensor   # [b, Nv] int64 {0,1}
    global_text: torch.Tensor  # [b, 1, D] fp32
    global_video: torch.Tensor  # [b, 1, D] fp32
    idx: torch.Tensor          # [b] int64

    def to(self, device):
        return HeadInputs(*[getattr(self, f).to(device) for f in self.__dataclass_fields__])


def _ragged_mask(g, b, n, lo):
    lens = torch.randint(lo, n + 1, (b,), generator=g)
    return (torch.arange(n)[None, :] < lens[:, None]).to(torch.int64)


def make_batch(b, nt, nv, d=512, seed=1234, rank=0, sigma=6.0, ragged=True) -> HeadInputs:
    g = torch.Generator().manual_seed(seed + rank)
    base = torch.randn(b, 1, d, generator=g)
    text = base + sigma * torch.randn(b, nt, d, generator=g)
    video = base + sigma * torch.randn(b, nv, d, generator=g)
    if ragged:
        tmask = _ragged_mask(g, b, nt, min(3, nt))
        vmask = _ragged_mask(g, b, nv, max(1, nv // 3))
    else:
        tmask = torch.ones(b, nt, dtype=torch.int64)
        vmask = torch.ones(b, nv, dtype=torch.int64)
    gt = 3.0 * torch.nn.functional.normalize(base + 0.3 * torch.randn(b, 1, d, generator=g), dim=-1)
    gv = 3.0 * torch.nn.functional.normalize(base + 0.3 * torch.randn(b, 1, d, generator=g), dim=-1)
    idx = torch.arange(b, dtype=torch.int64) + rank * b
    return HeadInputs(text.contiguous(), video.contiguous(), tmask, vmask, gt, gv, idx)


def make_bank(m, nt, nv, d=512, seed=999, sigma=6.0):
    """Memory bank rows (all-ones masks), SURVEY.md §8(d)."""
    h = make_batch(m, nt, nv, d=d, seed=seed, rank=0, sigma=sigma, ragged=False)
    return SimpleNamespace(mb_ind=h.idx, mb_feat_t=h.text_feat, mb_feat_v=h.video_feat,
                           mb_mask_t=h.text_mask, mb_mask_v=h.video_mask)


def make_mlp_params(d=512, seed=0, names=("text_weight_fc", "video_weight_fc",
                                          "text_weight_fc1", "video_weight_fc1")):
    """Token-weight MLP parameters N(0, 0.02^2), zero bias (reference modeling.py:648-659)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for n in names:
        out[n] = {
            "0.weight": 0.02 * torch.randn(2 * d, d, generator=g),
            "0.bias": torch.zeros(2 * d),
            "2.weight": 0.02 * torch.randn(1, 2 * d, generator=g),
            "2.bias": torch.zeros(1),
        }
    return out


def default_config(**kw):
    """Hyper-parameters the head reads from ``config`` (reference args_parser.py:26-40)."""
    c = dict(centrality_scale=0.3, beta=0.7, num_neighbors=20, temperature=3.0,
             uniform_weight=1.0, neighbor_weight=1.0, kl_weight=1.0,
             world_size=1, local_rank=0, rank=0)
    c.update(kw)
    return SimpleNamespace(**c)


# ---- multi-sentence test sets (several captions per video, reference evaluator.py:216-251) ------------------
MS_CASES = {
    # name: (videos, max captions per video, integer-valued scores with many ties, NaN/inf entries, seed)
    # ties only at V <= 16: the reference's torch.argsort (stable=False) keeps equal scores in column order
    # there and in an implementation-defined order above (introsort), see oracle/metrics.py
    "ties16": (16, 6, True, True, 11),
    "plain150": (150, 6, False, False, 12),
    "nan120": (120, 5, False, True, 13),
}


def make_multi_sentence_case(V, maxlen, ties, nonfinite, seed):
    """Seeded caption x video matrix [T, V] + cut-off points (index of the last caption of every video)."""
    rng = np.random.RandomState(seed)
    lens = rng.randint(1, maxlen + 1, size=V)
    T = int(lens.sum())
    tgt = np.repeat(np.arange(V), lens)
    if ties:
        sim = rng.randint(0, 5, size=(T, V)).astype(np.float32)
    else:
        sim = rng.randn(T, V).astype(np.float32)
        sim[np.arange(T), tgt] += 2.0
    if nonfinite:
        sim[rng.rand(T, V) < 0.01] = np.nan
        sim[rng.rand(T, V) < 0.01] = np.inf
        sim[rng.rand(T, V) < 0.01] = -np.inf
    return sim, (np.cumsum(lens) - 1).astype(np.int64)


# ---- memory-bank prefill (reference utils/memory_bank.py:80-229) ------------------------------------------------
class ToyEncoder(torch.nn.Module):
    """Stand-in for the CLIP towers in prefill tests: deterministic features from the loader tensors, exposed as
    ``get_text_video_feat`` like the reference model, plus the bank attributes the manager writes."""

    def __init__(self, d=8):
        super().__init__()
        self.d = d
        self.scale = torch.nn.Parameter(torch.tensor(0.5))
        self.mb_ind = torch.tensor([], dtype=torch.long)
        self.mb_feat_t = torch.empty((0, 0, 0))
        self.mb_feat_v = torch.empty((0, 0, 0))
        self.mb_mask_t = torch.empty((0, 0))
        self.mb_mask_v = torch.empty((0, 0))
        self.mb_batch = 0

    def get_text_video_feat(self, text_ids, text_mask, video, video_mask, shaped=False):
        ramp = torch.arange(1, self.d + 1, device=text_ids.device, dtype=torch.float32)
        text = torch.sin(text_ids.float().unsqueeze(-1) * ramp * self.scale)
        vid = torch.cos(video.float().mean(dim=-1, keepdim=True) * ramp)
        return text.float(), vid.float()


def make_prefill_loader(n_batches, b, nt=4, nv=3, seed=77, rank=0):
    """List of loader batches ``(text_ids, text_mask, video, video_mask, inds, idx)`` (reference
    dataloader_retrieval.py:343) with ragged prefix masks and dataset indices shaped [b, 1]."""
    g = torch.Generator().manual_seed(seed + 1000 * rank)
    out = []
    for k in range(n_batches):
        ids = torch.randint(0, 400, (b, nt), generator=g)
        video = torch.randn(b, nv, 5, generator=g)
        tl = torch.randint(1, nt + 1, (b,), generator=g)
        vl = torch.randint(1, nv + 1, (b,), generator=g)
        tm = (torch.arange(nt)[None, :] < tl[:, None]).long()
        vm = (torch.arange(nv)[None, :] < vl[:, None]).long()
        inds = (torch.arange(b) + (k + 10 * rank) * b).view(b, 1)
        out.append((ids, tm, video, vm, inds, inds.clone()))
    return out
