"""Evaluation ranking: RetrievalMetrics.compute_metrics with the reference's return dict
(reference NeighborRetr/utils/metrics.py:38-79), computed by the rank-count kernel.

The reference sorts every row of -S on the host and locates the diagonal value in the sorted row.  The same
ranks follow from two counts per row — g = #{j: S[i,j] > S[i,i]}, e = #{j: S[i,j] == S[i,i]} — with the
reference's tie behaviour reproduced exactly: a row contributes the ranks g, g+1, ..., g+e-1.
Comparison-only, so ranks are bit-exact for the same fp32 matrix.
"""
from __future__ import annotations

from typing import Dict

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


class RetrievalMetrics:
    """Drop-in for the reference class's static ``compute_metrics`` (the tracking/printing helpers of the
    reference class are logging-only and stay in the reference)."""

    def __init__(self, logger=None):
        self.logger = logger

    @staticmethod
    def compute_metrics(similarity_matrix) -> Dict[str, float]:
        """similarity_matrix: np.ndarray [N,N] (as the reference passes it) or a CUDA tensor."""
        if not torch.cuda.is_available():
            raise RuntimeError("RetrievalMetrics.compute_metrics: no CUDA device (no CPU fallback exists)")
        if isinstance(similarity_matrix, np.ndarray):
            s = torch.from_numpy(np.ascontiguousarray(similarity_matrix, dtype=np.float32)).cuda()
        else:
            s = similarity_matrix
        if s.dim() != 2 or s.shape[0] > s.shape[1]:
            raise ValueError(f"compute_metrics: need [Q,N] with a diagonal, got {tuple(s.shape)}")
        gt, eq = ops.rank_counts(s)
        return metrics_from_counts(gt.cpu().numpy(), eq.cpu().numpy())
