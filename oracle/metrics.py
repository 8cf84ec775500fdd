"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/head.py header).

numpy restatement of ``RetrievalMetrics.compute_metrics`` (reference
NeighborRetr/utils/metrics.py:38-79): full sort of every row of -S, positions where the sorted row
equals the diagonal element, R@1/5/10/50, median and mean rank.  Comparison-only => integer-exact.

Ties: a row with g entries strictly greater than S[i,i] and e entries equal to it
(the diagonal included) contributes the ranks g, g+1, ..., g+e-1 — ``np.where`` lists every
position whose sorted value equals the diagonal — so ``len(cols)`` grows by e-1.
"""
import numpy as np


def compute_metrics(sim: np.ndarray) -> dict:
    sx = np.sort(-sim, axis=1)                      # metrics.py:58
    d = np.diag(-sim)[:, np.newaxis]                # :60-61
    ind = np.where((sx - d) == 0)[1]                # :63-66
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


def ranks_by_counting(sim: np.ndarray):
    """The counting form the CUDA kernel implements: per row g = #{j: S[i,j] > S[i,i]},
    e = #{j: S[i,j] == S[i,i]} (>= 1).  Returns (g, e) int64 arrays; ``cols`` is the
    concatenation over rows of range(g_i, g_i + e_i)."""
    d = np.diag(sim)[:, None]
    return (sim > d).sum(axis=1).astype(np.int64), (sim == d).sum(axis=1).astype(np.int64)
