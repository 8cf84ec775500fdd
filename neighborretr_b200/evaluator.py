"""Evaluation similarity matrix: _run_on_single_gpu with the reference's signature and return value
(reference NeighborRetr/training/evaluator.py:21-63).

The reference walks 64x64 tiles, runs local_level (including the token-weight MLPs) per tile pair and copies
every tile to the host.  Tiling does not change values, so here the token weights are
evaluated once per modality, the whole [Nq,Ng] matrix comes from one pair of max-sim launches, and there is a
single device->host copy.  ``mini_batch`` is accepted for signature compatibility and bounds the MLP batch.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .modeling import _token_weights


def _eval_operands(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, mini_batch):
    """(flattened masks, token weights of both modalities, head precision): the MLPs run once per modality."""
    t_mask = t_mask_list.view(-1, t_mask_list.shape[-1])
    v_mask = v_mask_list.view(-1, v_mask_list.shape[-1])
    chunk = max(int(mini_batch), 1) * 64
    prec = model._head_precision() if hasattr(model, "_head_precision") else "fp32"
    lowp = model._mlp_precision() if hasattr(model, "_mlp_precision") else ("tf32" if prec == "bf16" else "fp32")
    tw = torch.cat([_token_weights(model.text_weight_fc, f, m, lowp)
                    for f, m in zip(torch.split(t_feat_list, chunk), torch.split(t_mask, chunk))])
    vw = torch.cat([_token_weights(model.video_weight_fc, f, m, lowp)
                    for f, m in zip(torch.split(v_feat_list, chunk), torch.split(v_mask, chunk))])
    return t_mask, v_mask, tw, vw, prec


def similarity_matrix(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, mini_batch=64):
    """CUDA tensor [Nq, Ng] of local_level similarities."""
    with torch.no_grad():
        t_mask, v_mask, tw, vw, prec = _eval_operands(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list,
                                                      mini_batch)
        s, _ = ops.maxsim(t_feat_list, v_feat_list, tw, vw, t_mask, v_mask, prec)
    return s


def retrieval_counts_fused(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, mini_batch=64, video0=0,
                           total=None, reduce_diag=None):
    """Rank counts of both retrieval directions WITHOUT the similarity matrix (north_star: the B x B logits never
    round-trip HBM; at a 100 000 gallery the fp32 matrix of reference training/evaluator.py:21-63 is 37 GiB): the
    texts are all `total` queries, the videos the gallery shard starting at pair id `video0`.  Pass 1 contracts only
    the tiles that hold a positive pair and writes its score; `reduce_diag` (sharded galleries: a SUM all-reduce)
    completes that vector; pass 2 contracts the whole block and compares every accumulator value with the
    positives' scores in the kernel epilogue.  Returns int32 CUDA vectors (gt_t, eq_t, gt_v, eq_v): per text over
    this shard's videos, per video of the shard over all texts — the counts ops.rank_counts gives on the block and on
    its transpose, from which metrics_from_counts reproduces compute_metrics (reference utils/metrics.py:38-79)."""
    with torch.no_grad():
        t_mask, v_mask, tw, vw, prec = _eval_operands(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list,
                                                      mini_batch)
        if prec == "fp32":
            raise RuntimeError("retrieval_counts_fused: the fused ranks come from the tensor-core kernel; use "
                               "head_precision 'bf16' or 'bf16x3' (or the materialised path for CUDA-core fp32)")
        r = ops.FusedRanker(t_feat_list, v_feat_list, tw, vw, t_mask, v_mask, prec, text0=0, video0=video0)
        n_pairs = int(total) if total is not None else max(t_feat_list.shape[0], video0 + v_feat_list.shape[0])
        diag = r.diagonal(n_pairs)
        if reduce_diag is not None:
            reduce_diag(diag)
        return r.counts(diag)


def retrieval_metrics_fused(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, mini_batch=64):
    """(text->video, video->text) metric dicts of a square test set on one GPU, no similarity matrix: what
    compute_metrics(sim_matrix), compute_metrics(sim_matrix.T) return after _run_on_single_gpu (reference
    training/evaluator.py:253-266)."""
    from .metrics import metrics_from_counts
    if t_feat_list.shape[0] != v_feat_list.shape[0]:
        raise ValueError("retrieval_metrics_fused: compute_metrics needs a square query x gallery problem")
    cnt = retrieval_counts_fused(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, mini_batch)
    c = torch.stack(cnt).cpu().numpy()                  # one device->host copy
    return metrics_from_counts(c[0], c[1]), metrics_from_counts(c[2], c[3])


def _run_on_single_gpu(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, mini_batch=64):
    s = similarity_matrix(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, mini_batch)
    sim_matrix = s.cpu().numpy()
    return sim_matrix, sim_matrix.T


def sharded_retrieval(model, t_mask_list, v_mask_list, t_feat_list, v_feat_list, topk=10, fused=False):
    """Column-sharded evaluation over the default process group: every rank has all features
    (as after the reference's gather, training/evaluator.py:173-189), computes only S[:, cols_r] for its shard of
    the video gallery and never materialises the full matrix.  Exact global ranks come from per-shard counts
    (all_reduce of int32 vectors) against the positive's score; top-k lists are merged across shards
    (ties -> lower global index).  Returns (t2v metrics, v2t metrics, (topk values, topk video ids) [Q,k]).
    fused=True: the counts come straight from the contraction's epilogue (retrieval_counts_fused) — not even the
    shard's [Q, N/W] block exists; the top-k lists are then not produced (third value None)."""
    import torch.distributed as dist
    from .metrics import metrics_from_counts
    W, r = dist.get_world_size(), dist.get_rank()
    Q, N = t_feat_list.shape[0], v_feat_list.shape[0]
    if Q != N:
        raise ValueError("sharded_retrieval: compute_metrics needs a square query x gallery problem")
    n = (N + W - 1) // W
    c0, c1 = min(r * n, N), min((r + 1) * n, N)
    if fused:
        dev = t_feat_list.device
        cnt = torch.zeros(2, Q, dtype=torch.int32, device=dev)
        cnt_v = torch.zeros(2, W * n, dtype=torch.int32, device=dev)
        if c1 > c0:
            gt_t, eq_t, gt_v, eq_v = retrieval_counts_fused(
                model, t_mask_list, v_mask_list[c0:c1], t_feat_list, v_feat_list[c0:c1], video0=c0, total=Q,
                reduce_diag=lambda d: dist.all_reduce(d, op=dist.ReduceOp.SUM))
            cnt[0], cnt[1] = gt_t, eq_t
            cnt_v[0, c0:c1], cnt_v[1, c0:c1] = gt_v, eq_v
        else:                                   # an empty shard still takes part in the collectives
            dist.all_reduce(torch.zeros(Q, dtype=torch.float32, device=dev), op=dist.ReduceOp.SUM)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(cnt_v, op=dist.ReduceOp.SUM)
        c = cnt.cpu().numpy(); cv = cnt_v[:, :N].cpu().numpy()
        return metrics_from_counts(c[0], c[1]), metrics_from_counts(cv[0], cv[1]), None
    s = similarity_matrix(model, t_mask_list, v_mask_list[c0:c1], t_feat_list, v_feat_list[c0:c1])   # [Q, c1-c0]
    dev = s.device
    # positives' scores: owned by the rank whose shard holds column q
    diag = torch.zeros(Q, dtype=torch.float32, device=dev)
    if c1 > c0:
        diag[c0:c1] = s[c0:c1].diagonal()
    dist.all_reduce(diag, op=dist.ReduceOp.SUM)
    cnt = torch.zeros(2, Q, dtype=torch.int32, device=dev)
    if c1 > c0:
        ops.rank_counts(s, diag=diag, gt=cnt[0], eq=cnt[1])
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    # video -> text: rows of S^T for this rank's videos are complete locally
    cnt_v = torch.zeros(2, W * n, dtype=torch.int32, device=dev)
    if c1 > c0:
        st = s.t().contiguous()                                       # [c1-c0, Q]
        ops.rank_counts(st, diag_col0=c0, gt=cnt_v[0, c0:c1], eq=cnt_v[1, c0:c1])
    dist.all_reduce(cnt_v, op=dist.ReduceOp.SUM)
    # top-k video ids per text query
    k = min(topk, n)
    vals = torch.full((Q, k), float("-inf"), device=dev)
    idx = torch.full((Q, k), -1, dtype=torch.int32, device=dev)
    if c1 > c0:
        kk = min(k, c1 - c0)
        v_, i_ = ops.topk_rows(s, kk, col_offset=c0)
        vals[:, :kk], idx[:, :kk] = v_, i_
    allv = torch.empty(W * Q, k, device=dev); alli = torch.empty(W * Q, k, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(allv, vals)
    dist.all_gather_into_tensor(alli, idx)
    tv, ti = ops.topk_merge(allv.view(W, Q, k), alli.view(W, Q, k))
    c = cnt.cpu().numpy(); cv = cnt_v[:, :N].cpu().numpy()
    return metrics_from_counts(c[0], c[1]), metrics_from_counts(cv[0], cv[1]), (tv, ti)


def multi_sentence_metrics(sim_matrix, cut_off_points):
    """Multi-sentence branch of eval_epoch (reference training/evaluator.py:216-251): sim_matrix [T, V] caption x
    video similarities (np.ndarray as returned by _run_on_single_gpu, or a CUDA tensor), cut_off_points[i] = index
    of the last caption of video i.  Returns (text->video metrics, video->text metrics) with the reference's keys.

    The reference pads the matrix to [V, maxlen, V] with -inf rows on the host, double-argsorts it and takes a
    max over the padded axis; the same numbers come from the un-padded matrix in two passes over it — the rank
    of every caption's own video (target-rank kernel) and the per-video max over its captions (group-max kernel,
    written transposed) followed by the ordinary rank count."""
    from . import metrics as mt
    s = mt._to_cuda_f32(sim_matrix)
    if s.dim() != 2:
        raise ValueError(f"multi_sentence_metrics: need [T, V], got {tuple(s.shape)}")
    T, V = s.shape
    group_start, target = mt.group_layout(cut_off_points, total_rows=T)
    if len(group_start) - 1 != V:
        raise ValueError(f"multi_sentence_metrics: {len(group_start) - 1} cut-off points for {V} videos")
    ranks, valid = mt.multi_sentence_ranks(s, torch.from_numpy(target).cuda())
    order = mt.slot_major_order(group_start, target)
    tv = mt.multi_sentence_rank_scalars(ranks[order][valid[order]])
    v2t = ops.group_max_t(s, torch.from_numpy(group_start).cuda())              # [V videos, V caption groups]
    vt = mt.RetrievalMetrics.compute_metrics(v2t)
    return tv, vt


def gather_eval_features(args, ids_t, mask_t, mask_v, feat_t, feat_v):
    """Gather + reorder-by-id step of eval_epoch (reference training/evaluator.py:173-189): every rank holds a
    shard of the test set with the dataset indices ``ids_t``; after the gather row ids_t[k] of each tensor is the
    k-th gathered row, and the tensors are trimmed to max(id)+1 rows.  Pure data movement: contiguous
    all_gather_into_tensor (until_module.AllGather) + one index_copy per tensor."""
    from .until_module import AllGather
    ids = AllGather.apply(ids_t, args).reshape(-1)
    n = int(ids.max().item()) + 1
    out = []
    for t in (mask_t, mask_v, feat_t, feat_v):
        g = AllGather.apply(t, args)
        out.append(g.index_copy(0, ids, g)[:n])          # g[ids] = g.clone()  (evaluator.py:180-183)
    return (ids, *out)
