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


def shard_counts(sim: np.ndarray, lo: int, hi: int):
    """What one gallery shard contributes to the ranks of a square test set (the host logic of
    neighborretr_b200.evaluator.sharded_retrieval and of the fused rank pass nr_maxsim2_rank): from the column block
    sim[:, lo:hi] and the positives' scores d = diag(sim),
      per text q   : (#{j in shard: S[q,j] > d[q]}, #{j in shard: S[q,j] == d[q]})   -> summed over shards = ranks_by_counting(sim)
      per video v in the shard: (#{q: S[q,v] > d[v]}, #{q: S[q,v] == d[v]})           -> complete = ranks_by_counting(sim.T)[lo:hi]
    Returns (gt_t, eq_t, gt_v, eq_v) int64 arrays."""
    d = np.diag(sim)
    blk = sim[:, lo:hi]
    return ((blk > d[:, None]).sum(axis=1).astype(np.int64), (blk == d[:, None]).sum(axis=1).astype(np.int64),
            (blk > d[None, lo:hi]).sum(axis=0).astype(np.int64), (blk == d[None, lo:hi]).sum(axis=0).astype(np.int64))


# ---- multi-sentence test sets (several captions per video; MSVD) ---------------------------------------------

def multi_sentence_reshape(sim: np.ndarray, cut_off_points) -> np.ndarray:
    """[T, V] caption x video matrix -> [V, maxlen, V], video i's captions in slab i, missing caption slots
    filled with -inf rows (reference training/evaluator.py:216-239; ``cut_off_points`` = index of the last
    caption of every video, captions of a video contiguous)."""
    ends = [int(c) + 1 for c in cut_off_points]                     # evaluator.py:227
    starts = [0] + ends[:-1]
    maxlen = max(e - s for s, e in zip(starts, ends))               # :228
    slabs = []
    for s, e in zip(starts, ends):                                  # :232-238
        pad = np.full((maxlen - (e - s), sim.shape[1]), -np.inf, dtype=sim.dtype)
        slabs.append(np.concatenate((sim[s:e], pad), axis=0))
    return np.stack(slabs, axis=0)


def tensor_text_to_video_metrics(sim_tensor: np.ndarray, top_k=(1, 5, 10, 50)) -> dict:
    """Reference utils/metrics.py:81-122 on the padded [V, maxlen, V] tensor: rank of video i in the row of its
    l-th caption = position of column i in the descending (stable) argsort; caption slots whose own score is
    +-inf/NaN are dropped.  Scalars reproduce the reference's types: R@k is an int64 tensor divided by an int
    (float32 division), MedianR is torch.median = LOWER median, MeanR/Std_Rank are float64 numpy."""
    x = np.transpose(np.asarray(sim_tensor), (1, 0, 2))             # :96  [maxlen, V(group), V(video)]
    key = np.where(np.isnan(x), np.inf, x)                          # torch sorts NaN as the largest value
    first = np.argsort(-key, axis=-1, kind="stable")                # :97
    second = np.argsort(first, axis=-1, kind="stable")              # :98
    L, V, _ = x.shape
    ranks = second[:, np.arange(V), np.arange(V)].reshape(-1)       # :101
    own = x[:, np.arange(V), np.arange(V)].reshape(-1)              # :104 (same (l, i) order)
    valid = ranks[~(np.isinf(own) | np.isnan(own))].astype(np.int64)   # :105-106
    n = len(valid)
    res = {f"R{k}": float(np.float32(int(np.sum(valid < k)) * 100) / np.float32(n)) for k in top_k}   # :112
    res["MedianR"] = float(np.sort(valid + 1)[(n - 1) // 2])        # :113 torch.median -> lower of the two middles
    res["MeanR"] = float(np.mean(valid + 1))                        # :114
    res["Std_Rank"] = float(np.std(valid + 1))                      # :115
    res["MR"] = res["MedianR"]                                      # :116
    return res, valid


def tensor_video_to_text_sim(sim_tensor: np.ndarray) -> np.ndarray:
    """Reference utils/metrics.py:124-145: NaN -> -inf, max over the caption slots of every slab, squeeze,
    transpose: out[j, i] = max_l sim_tensor[i, l, j]  ([video j, caption group i])."""
    x = np.array(sim_tensor, copy=True)
    x[x != x] = -np.inf                                             # :139
    return np.max(x, axis=1).T                                      # :142-144


def multi_sentence_ranks_by_counting(sim: np.ndarray, cut_off_points):
    """The counting form the CUDA kernel implements on the UN-padded [T, V] matrix: for caption t of video c,
    rank = #{j: S[t,j] > S[t,c] or NaN} + #{j < c: S[t,j] == S[t,c]}; valid unless S[t,c] is +-inf/NaN.
    Returns (ranks int64 [T], valid bool [T]) in caption order."""
    ends = np.asarray(cut_off_points, dtype=np.int64) + 1
    lens = np.diff(np.concatenate([[0], ends]))
    tgt = np.repeat(np.arange(len(ends)), lens)
    T = sim.shape[0]
    sd = sim[np.arange(T), tgt][:, None]
    gt = ((sim > sd) | np.isnan(sim)).sum(axis=1)
    before = np.arange(sim.shape[1])[None, :] < tgt[:, None]
    eqb = ((sim == sd) & before).sum(axis=1)
    valid = ~(np.isinf(sd[:, 0]) | np.isnan(sd[:, 0]))
    return (gt + eqb).astype(np.int64), valid
