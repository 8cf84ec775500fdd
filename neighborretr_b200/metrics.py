"""Evaluation ranking: RetrievalMetrics.compute_metrics with the reference's return dict
(reference NeighborRetr/utils/metrics.py:38-79), computed by the rank-count kernel, and the multi-sentence
variants tensor_text_to_video_metrics / tensor_video_to_text_sim (:81-145) computed by the target-rank and
group-max kernels (csrc/multisent.cu).

The reference sorts every row of -S on the host and locates the diagonal value in the sorted row.  The same
ranks follow from two counts per row — g = #{j: S[i,j] > S[i,i]}, e = #{j: S[i,j] == S[i,i]} — with the
reference's tie behaviour reproduced exactly: a row contributes the ranks g, g+1, ..., g+e-1.
Comparison-only, so ranks are bit-exact for the same fp32 matrix.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from . import ops


def metrics_from_counts(gt: np.ndarray, eq: np.ndarray) -> Dict[str, float]:
    """Host-side finish: expand (g, e) into the reference's `ind` vector and reduce."""
    gt = np.asarray(gt, dtype=np.int64)
    eq = np.asarray(eq, dtype=np.int64)
    if np.all(eq == 1):
        ind = gt
    else:
        ind = np.concatenate([np.arange(g, g + e) for g, e in zip(gt, eq)]) if len(gt) else gt
    m = {}
    m["R1"] = float(np.sum(ind == 0)) * 100 / len(ind)
    m["R5"] = float(np.sum(ind < 5)) * 100 / len(ind)
    m["R10"] = float(np.sum(ind < 10)) * 100 / len(ind)
    m["R50"] = float(np.sum(ind < 50)) * 100 / len(ind)
    m["MR"] = float(np.median(ind)) + 1
    m["MedianR"] = m["MR"]
    m["MeanR"] = float(np.mean(ind)) + 1
    m["cols"] = [int(i) for i in list(ind)]
    return m


def _to_cuda_f32(x):
    """np.ndarray / CPU tensor (what the reference's callers pass) or CUDA tensor -> contiguous CUDA fp32."""
    if not torch.cuda.is_available():
        raise RuntimeError("neighborretr_b200.metrics: no CUDA device (no CPU fallback exists)")
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
    if not torch.is_tensor(x):
        raise TypeError(f"expected np.ndarray or torch.Tensor, got {type(x).__name__}")
    return x.detach().to(device="cuda", dtype=torch.float32).contiguous()


def group_layout(cut_off_points: Sequence[int], total_rows=None):
    """cut_off_points (index of the last caption of every video, reference evaluator.py:227-228) ->
    (group_start int32 [G+1], target int32 [T]) as numpy arrays.  Captions of a video are contiguous."""
    ends = np.asarray([int(c) + 1 for c in cut_off_points], dtype=np.int64)
    if ends.ndim != 1 or len(ends) == 0:
        raise ValueError("group_layout: need at least one cut-off point")
    starts = np.concatenate([[0], ends])
    lens = np.diff(starts)
    if np.any(lens < 0):
        raise ValueError("group_layout: cut-off points must be non-decreasing")
    if total_rows is not None and int(ends[-1]) != int(total_rows):
        raise ValueError(f"group_layout: last cut-off point {int(ends[-1]) - 1} does not end the {total_rows} rows")
    target = np.repeat(np.arange(len(ends), dtype=np.int32), lens)
    return starts.astype(np.int32), target


def slot_major_order(group_start: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Permutation of the captions into the order the reference flattens its ranks in (metrics.py:96-101: the
    l-th captions of all videos, then the (l+1)-th, ...).  Only Std_Rank depends on it (float64 summation order)."""
    slot = np.arange(len(target)) - np.asarray(group_start, dtype=np.int64)[target]
    return np.lexsort((target, slot))


def multi_sentence_rank_scalars(valid_ranks: np.ndarray, top_k=(1, 5, 10, 50)) -> Dict[str, float]:
    """Host-side finish of tensor_text_to_video_metrics with the reference's scalar types (metrics.py:112-116):
    R@k is an int64 tensor divided by an int (float32 division), MedianR = torch.median (LOWER median),
    MeanR / Std_Rank are float64 numpy reductions."""
    r = np.asarray(valid_ranks, dtype=np.int64)
    n = len(r)
    res = {f"R{k}": float(np.float32(int(np.sum(r < k)) * 100) / np.float32(n)) for k in top_k}
    res["MedianR"] = float(np.sort(r + 1)[(n - 1) // 2])
    res["MeanR"] = float(np.mean(r + 1))
    res["Std_Rank"] = float(np.std(r + 1))
    res["MR"] = res["MedianR"]
    return res


def multi_sentence_ranks(sim, target):
    """Ranks of every caption's own video from the UN-padded caption x video matrix: sim [T,V] (np / CPU / CUDA),
    target [T] = video column of each caption.  Returns (ranks int64 [T], valid bool [T]) numpy arrays."""
    s = _to_cuda_f32(sim)
    if s.dim() != 2:
        raise ValueError(f"multi_sentence_ranks: need [T,V], got {tuple(s.shape)}")
    t = torch.as_tensor(np.asarray(target, dtype=np.int32)).cuda() if not torch.is_tensor(target) else target.cuda()
    gt, eqb, valid = ops.rank_counts_target(s, t)
    packed = torch.stack([gt, eqb, valid]).cpu().numpy()          # one device->host copy
    return (packed[0].astype(np.int64) + packed[1]), packed[2].astype(bool)


class RetrievalMetrics:
    """Drop-in for the reference class's static ``compute_metrics``, ``tensor_text_to_video_metrics`` and
    ``tensor_video_to_text_sim`` (the tracking/printing helpers of the reference class are logging-only and stay
    in the reference)."""

    def __init__(self, logger=None):
        self.logger = logger

    @staticmethod
    def compute_metrics(similarity_matrix) -> Dict[str, float]:
        """similarity_matrix: np.ndarray [N,N] (as the reference passes it) or a CUDA tensor."""
        s = _to_cuda_f32(similarity_matrix)
        if s.dim() != 2 or s.shape[0] > s.shape[1]:
            raise ValueError(f"compute_metrics: need [Q,N] with a diagonal, got {tuple(s.shape)}")
        gt, eq = ops.rank_counts(s)
        # one device->host copy: counts + the positives' scores (bit pattern)
        packed = torch.stack([gt, eq, s.diagonal().contiguous().view(torch.int32)]).cpu().numpy()
        gt, eq, d = packed[0], packed[1], packed[2].view(np.float32)
        # a positive whose score is +-inf never matches in the reference (inf - inf = NaN != 0 at metrics.py:63-64)
        # and drops out of `cols` like a NaN one; such scores only occur in the -inf padded multi-sentence path
        if not np.all(np.isfinite(d)):
            eq = np.where(np.isfinite(d), eq, 0)
        return metrics_from_counts(gt, eq)

    @staticmethod
    def tensor_text_to_video_metrics(sim_tensor, top_k: List[int] = [1, 5, 10, 50]) -> Dict[str, float]:
        """Reference metrics.py:81-122.  sim_tensor: the -inf padded [V, maxlen, V] tensor eval_epoch builds
        (slab i = captions of video i, evaluator.py:232-238), np.ndarray or tensor.  Row (i, l) is ranked against
        column i by the target-rank kernel; padded slots (own score -inf) are dropped like in the reference.
        Equal scores: the reference's ``torch.argsort(stable=False)`` leaves their order implementation-defined
        (column order up to 16 columns, introsort above); here it is always column order."""
        s = _to_cuda_f32(sim_tensor)
        if s.dim() != 3 or s.shape[0] != s.shape[2]:
            raise ValueError(f"tensor_text_to_video_metrics: need [V, maxlen, V], got {tuple(s.shape)}")
        V, L, _ = s.shape
        target = torch.arange(V, dtype=torch.int32, device=s.device).repeat_interleave(L)
        ranks, valid = multi_sentence_ranks(s.view(V * L, V), target)
        ranks, valid = ranks.reshape(V, L).T.reshape(-1), valid.reshape(V, L).T.reshape(-1)   # (slot, video) order
        return multi_sentence_rank_scalars(ranks[valid], top_k)

    @staticmethod
    def tensor_video_to_text_sim(sim_tensor) -> torch.Tensor:
        """Reference metrics.py:124-145: [V, maxlen, V'] padded tensor -> CPU tensor [V', V] with
        out[j, i] = max_l sim_tensor[i, l, j], NaN read as -inf.  (The reference also overwrites the NaNs of a
        tensor argument in place; the argument is left untouched here.)"""
        s = _to_cuda_f32(sim_tensor)
        if s.dim() != 3:
            raise ValueError(f"tensor_video_to_text_sim: need a 3-D tensor, got {tuple(s.shape)}")
        V, L, N = s.shape
        gs = torch.arange(0, (V + 1) * L, L, dtype=torch.int32, device=s.device)
        return ops.group_max_t(s.view(V * L, N), gs).cpu()
