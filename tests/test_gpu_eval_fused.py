"""Evaluation ranks straight from the accumulator (nr_maxsim2_rank, evaluator.retrieval_counts_fused) against the
materialised path: similarity matrix (evaluator.similarity_matrix) + rank-count kernel (ops.rank_counts), which is
itself pinned to the reference's compute_metrics golden vectors (tests/test_gpu_parity.py, reference
utils/metrics.py:38-79) and to the oracle (oracle/metrics.py).  Integer work: the counts must be EQUAL — the fused
epilogue recomputes the very same fp32 similarities (same operands, same accumulation order per element)."""
import numpy as np
import pytest
import torch

from helpers import make_head
from neighborretr_b200 import ops, synth
from neighborretr_b200.evaluator import (_eval_operands, retrieval_counts_fused, retrieval_metrics_fused,
                                         similarity_matrix)
from neighborretr_b200.metrics import RetrievalMetrics, metrics_from_counts
from oracle import metrics as OM

pytestmark = pytest.mark.gpu


def _case(q, nt, nv, d, prec, seed=5, ties=False):
    """Model, inputs and ONE evaluation of the token weights shared by both paths: the weight MLP's second layer is
    accumulated with float atomics (gemm_tc.cu), so two evaluations of it differ in the last bit and would move
    near-ties between two otherwise identical computations."""
    cfg = synth.default_config()
    m = make_head(d, cfg, synth.make_mlp_params(d=d), prec).eval()
    h = synth.make_batch(q, nt, nv, d=d, seed=seed).to("cuda")
    with torch.no_grad():
        tm, vm, tw, vw, _ = _eval_operands(m, h.text_mask, h.video_mask, h.text_feat, h.video_feat, 64)
    if ties:          # two identical gallery videos / two identical queries: exact ties with the positive
        h.video_feat[5], vm[5], vw[5] = h.video_feat[3], vm[3], vw[3]
        h.text_feat[9], tm[9], tw[9] = h.text_feat[7], tm[7], tw[7]
    return m, h, tm, vm, tw, vw


def _materialised_counts(S, lo, hi):
    """Counts of the block S[:, lo:hi] and of its transpose against the positives' scores S[q, q]."""
    diag = S.diagonal().contiguous()
    blk = S[:, lo:hi].contiguous()
    gt_t, eq_t = ops.rank_counts(blk, diag=diag)
    gt_v, eq_v = ops.rank_counts(blk.t().contiguous(), diag=diag[lo:hi].contiguous())
    return gt_t, eq_t, gt_v, eq_v


@pytest.mark.parametrize("q,nt,nv,d,prec", [
    (101, 24, 12, 512, "bf16"),          # MSR-VTT tokens, ragged last tiles on both sides
    (101, 24, 12, 512, "bf16x3"),        # split operands (K = 3d), exact-order keys
    (1000, 24, 12, 512, "bf16"),         # the reference's own test-set size: several waves of tiles
    (37, 64, 64, 512, "bf16"),           # ActivityNet-shaped tokens: 2 x 4 samples per tile
    (45, 24, 20, 512, "bf16"),           # video plays X (Ny = 20 has no instantiation): swapped roles
    (3, 24, 12, 512, "bf16"),            # fewer samples than one tile
])
def test_fused_counts_equal_materialised(q, nt, nv, d, prec):
    m, h, tm, vm, tw, vw = _case(q, nt, nv, d, prec, ties=q > 10)
    with torch.no_grad():
        S, _ = ops.maxsim(h.text_feat, h.video_feat, tw, vw, tm, vm, prec)
    want = _materialised_counts(S, 0, q)
    r = ops.FusedRanker(h.text_feat, h.video_feat, tw, vw, tm, vm, prec)
    diag = r.diagonal(q)
    assert torch.equal(diag, S.diagonal())                            # pass 1 = the matrix's own diagonal, bit for bit
    got = r.counts(diag)
    for g, w, name in zip(got, want, ("gt_t", "eq_t", "gt_v", "eq_v")):
        assert torch.equal(g, w), (name, (g != w).sum().item())
    if q > 10:
        assert int(got[1][3]) == 2 and int(got[3][7]) == 2          # the planted ties are seen as ties
    # the counts give the metric dicts of compute_metrics on the matrix / its transpose = the oracle's sort-based ranks
    c = torch.stack(got).cpu().numpy()
    tv, vt = metrics_from_counts(c[0], c[1]), metrics_from_counts(c[2], c[3])
    assert tv == RetrievalMetrics.compute_metrics(S)
    assert vt == RetrievalMetrics.compute_metrics(S.t().contiguous())
    assert tv["cols"] == [int(v) for v in OM.compute_metrics(S.cpu().numpy())["cols"]]
    assert vt["cols"] == [int(v) for v in OM.compute_metrics(S.t().cpu().numpy())["cols"]]


@pytest.mark.parametrize("lo,hi", [(0, 40), (40, 77), (77, 101), (95, 101)])
def test_fused_counts_of_a_gallery_shard(lo, hi):
    """Column shard [lo, hi) of the gallery with all queries: positives of the shard from pass 1, the rest of the
    diagonal supplied by the caller (the all-reduce of the sharded evaluation)."""
    q, prec = 101, "bf16"
    m, h, tm, vm, tw, vw = _case(q, 24, 12, 512, prec, seed=11, ties=True)
    with torch.no_grad():
        S, _ = ops.maxsim(h.text_feat, h.video_feat, tw, vw, tm, vm, prec)
    full_diag = S.diagonal().contiguous()
    r = ops.FusedRanker(h.text_feat, h.video_feat[lo:hi], tw, vw[lo:hi], tm, vm[lo:hi], prec, text0=0, video0=lo)
    own = r.diagonal(q)
    assert torch.equal(own[lo:hi], full_diag[lo:hi])                  # the values pass 2 recomputes
    assert float(own[:lo].abs().sum()) == 0.0 and float(own[hi:].abs().sum()) == 0.0
    got = r.counts(full_diag)
    want = _materialised_counts(S, lo, hi)
    for g, w, name in zip(got, want, ("gt_t", "eq_t", "gt_v", "eq_v")):
        assert torch.equal(g, w), (name, (g != w).sum().item())
    # and the oracle's numpy statement of a shard's contribution (oracle/metrics.py: shard_counts) on the same matrix
    for g, w, name in zip(got, OM.shard_counts(S.cpu().numpy(), lo, hi), ("gt_t", "eq_t", "gt_v", "eq_v")):
        assert np.array_equal(g.cpu().numpy().astype(np.int64), w), name


def test_model_level_fused_metrics():
    """evaluator.retrieval_metrics_fused against compute_metrics(similarity_matrix): two separate evaluations of the
    token-weight MLPs (last-bit differences, see _case), so near-ties may move by one rank."""
    m, h, *_ = _case(1000, 24, 12, 512, "bf16")
    S = similarity_matrix(m, h.text_mask, h.video_mask, h.text_feat, h.video_feat)
    tv, vt = retrieval_metrics_fused(m, h.text_mask, h.video_mask, h.text_feat, h.video_feat)
    for got, want in ((tv, RetrievalMetrics.compute_metrics(S)), (vt, RetrievalMetrics.compute_metrics(S.t().contiguous()))):
        a, b = np.asarray(got["cols"]), np.asarray(want["cols"])
        if a.shape == b.shape:             # (an exact tie with a positive adds an entry: reference tie expansion)
            assert np.abs(a - b).max() <= 1 and (a != b).mean() < 0.01
        for k in ("R1", "R5", "R10", "R50", "MeanR"):
            assert abs(got[k] - want[k]) <= 0.2
    cnt = retrieval_counts_fused(m, h.text_mask, h.video_mask[100:300], h.text_feat, h.video_feat[100:300], video0=100,
                                 total=1000, reduce_diag=lambda d: d.copy_(S.diagonal()))
    assert [tuple(c.shape) for c in cnt] == [(1000,), (1000,), (200,), (200,)]


def test_fused_ranker_rejects_what_it_cannot_do():
    m, h, *_ = _case(8, 24, 12, 512, "fp32")
    with pytest.raises(RuntimeError):
        retrieval_counts_fused(m, h.text_mask, h.video_mask, h.text_feat, h.video_feat)
    m, h, *_ = _case(8, 24, 12, 512, "bf16")
    tw = torch.full((8, 24), 1 / 24, device="cuda"); vw = torch.full((8, 12), 1 / 12, device="cuda")
    r = ops.FusedRanker(h.text_feat, h.video_feat, tw, vw, h.text_mask, h.video_mask, "bf16")
    with pytest.raises(RuntimeError):
        r.counts(torch.zeros(4, device="cuda"))                      # diag shorter than the pair ids of the block
    with pytest.raises(RuntimeError):
        ops.FusedRanker(h.text_feat.cpu(), h.video_feat, tw, vw, h.text_mask, h.video_mask, "bf16")
