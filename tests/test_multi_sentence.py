"""Multi-sentence evaluation (several captions per video; reference utils/metrics.py:81-145 and the padding of
training/evaluator.py:216-251).

CPU part: the oracle restatement against golden vectors produced by the reference's own functions
(tests/golden/multi_sentence.npz, oracle/gen_golden.py), the counting form the kernels implement against the
padded double-argsort form, host bookkeeping, the eval gather/reorder under gloo (world 2).
GPU part: the kernels through the C ABI — bit-exact ranks / maxima against the golden vectors and the oracle."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import load_golden
from neighborretr_b200 import synth
from oracle import metrics as OM

TV_KEYS = ["R1", "R5", "R10", "R50", "MedianR", "MeanR", "Std_Rank", "MR"]
VT_KEYS = ["R1", "R5", "R10", "R50", "MR", "MedianR", "MeanR"]


def _vec(d, keys):
    return np.asarray([d[k] for k in keys], dtype=np.float64)


# ------------------------------------------------------------------------------------------------ CPU

@pytest.mark.parametrize("name", list(synth.MS_CASES))
def test_oracle_matches_reference_golden(name):
    gold = load_golden("multi_sentence")
    sim, cut = synth.make_multi_sentence_case(*synth.MS_CASES[name])
    pad = OM.multi_sentence_reshape(sim, cut)
    assert pad.shape[0] == pad.shape[2] == sim.shape[1]
    tv, valid = OM.tensor_text_to_video_metrics(pad)
    np.testing.assert_array_equal(_vec(tv, TV_KEYS), gold[f"{name}_tv"])
    v2t = OM.tensor_video_to_text_sim(pad)
    assert np.array_equal(v2t, gold[f"{name}_v2t_sim"], equal_nan=True)
    with np.errstate(invalid="ignore"):
        vt = OM.compute_metrics(v2t)
    assert np.array_equal(np.asarray(vt["cols"]), gold[f"{name}_vt_cols"])
    np.testing.assert_array_equal(_vec(vt, VT_KEYS), gold[f"{name}_vt"])


@pytest.mark.parametrize("name", list(synth.MS_CASES))
def test_counting_form_equals_padded_argsort_form(name):
    """rank = #greater(+NaN) + #equal-in-a-lower-column on the un-padded [T,V] matrix reproduces, caption by
    caption, the diagonal of the double argsort over the padded tensor; the per-group max over real captions
    equals the max over the padded slab."""
    sim, cut = synth.make_multi_sentence_case(*synth.MS_CASES[name])
    pad = OM.multi_sentence_reshape(sim, cut)
    _, valid_ranks = OM.tensor_text_to_video_metrics(pad)
    ranks, ok = OM.multi_sentence_ranks_by_counting(sim, cut)
    lens = np.diff(np.concatenate([[0], cut + 1]))
    starts = np.concatenate([[0], cut[:-1] + 1])
    order = np.asarray([starts[i] + l for l in range(pad.shape[1]) for i in range(len(cut)) if l < lens[i]])
    assert np.array_equal(ranks[order][ok[order]], valid_ranks)          # the reference's (slot, video) order
    from neighborretr_b200.metrics import group_layout, multi_sentence_rank_scalars, slot_major_order
    gs, tgt = group_layout(cut, total_rows=sim.shape[0])
    assert np.array_equal(slot_major_order(gs, tgt), order)
    assert gs.dtype == np.int32 and tgt.dtype == np.int32 and gs[-1] == sim.shape[0]
    x = np.where(np.isnan(sim), -np.inf, sim)
    gmax = np.stack([x[gs[i]:gs[i + 1]].max(axis=0) for i in range(len(cut))], axis=1)      # [V video, V group]
    assert np.array_equal(gmax, OM.tensor_video_to_text_sim(pad))
    tv, _ = OM.tensor_text_to_video_metrics(pad)
    assert multi_sentence_rank_scalars(ranks[order][ok[order]]) == tv


def test_group_layout_rejects_bad_cut_offs():
    from neighborretr_b200.metrics import group_layout
    gs, tgt = group_layout([0, 3, 5])
    assert gs.tolist() == [0, 1, 4, 6] and tgt.tolist() == [0, 1, 1, 1, 2, 2]
    with pytest.raises(ValueError):
        group_layout([3, 1])
    with pytest.raises(ValueError):
        group_layout([])
    with pytest.raises(ValueError):
        group_layout([0, 3, 5], total_rows=7)


def test_multi_sentence_surface_and_no_cpu_fallback():
    import inspect
    from neighborretr_b200 import evaluator as EV, metrics as MT
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(MT.RetrievalMetrics.tensor_text_to_video_metrics) == ["sim_tensor", "top_k"]
    assert sig(MT.RetrievalMetrics.tensor_video_to_text_sim) == ["sim_tensor"]
    assert inspect.signature(MT.RetrievalMetrics.tensor_text_to_video_metrics).parameters["top_k"].default == \
        [1, 5, 10, 50]
    assert sig(EV.multi_sentence_metrics) == ["sim_matrix", "cut_off_points"]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            MT.RetrievalMetrics.tensor_text_to_video_metrics(np.zeros((2, 1, 2), np.float32))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            MT.RetrievalMetrics.tensor_video_to_text_sim(torch.zeros(2, 1, 2))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            EV.multi_sentence_metrics(np.zeros((2, 2), np.float32), [0, 1])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from neighborretr_b200.evaluator import gather_eval_features
        args = SimpleNamespace(world_size=world, local_rank=rank)
        n, nt, nv, d = 10, 3, 2, 4                                    # test set of 10 items, 5 per rank, interleaved
        full_t = torch.arange(n * nt * d, dtype=torch.float32).reshape(n, nt, d)
        full_v = -torch.arange(n * nv * d, dtype=torch.float32).reshape(n, nv, d)
        full_mt = (torch.arange(n * nt).reshape(n, nt) % 3 > 0).long()
        full_mv = (torch.arange(n * nv).reshape(n, nv) % 2 > 0).long()
        mine = torch.tensor([9, 1, 4, 7, 2]) if rank == 0 else torch.tensor([0, 8, 3, 6, 5])
        ids, mt, mv, ft, fv = gather_eval_features(args, mine, full_mt[mine], full_mv[mine], full_t[mine],
                                                   full_v[mine])
        assert ids.tolist() == [9, 1, 4, 7, 2, 0, 8, 3, 6, 5]
        assert torch.equal(ft, full_t) and torch.equal(fv, full_v)
        assert torch.equal(mt, full_mt) and torch.equal(mv, full_mv)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gather_eval_features_world2_gloo():
    """evaluator.py:173-189: gather the ranks' shards and put row ids[k] <- gathered row k."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


# ------------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
@pytest.mark.parametrize("name", list(synth.MS_CASES))
def test_gpu_multi_sentence_matches_reference_golden(name):
    """Drop-in entry points on the padded tensor the reference's eval_epoch builds, and the un-padded fast path,
    against the reference's own outputs: every scalar, the video->text matrix and the v2t ranks bit-exact."""
    from neighborretr_b200.evaluator import multi_sentence_metrics
    from neighborretr_b200.metrics import RetrievalMetrics
    gold = load_golden("multi_sentence")
    sim, cut = synth.make_multi_sentence_case(*synth.MS_CASES[name])
    pad = OM.multi_sentence_reshape(sim, cut)
    tv = RetrievalMetrics.tensor_text_to_video_metrics(pad)
    np.testing.assert_array_equal(_vec(tv, TV_KEYS), gold[f"{name}_tv"])
    v2t = RetrievalMetrics.tensor_video_to_text_sim(pad)
    assert torch.is_tensor(v2t) and not v2t.is_cuda and tuple(v2t.shape) == (sim.shape[1], sim.shape[1])
    assert np.array_equal(v2t.numpy(), gold[f"{name}_v2t_sim"], equal_nan=True)
    vt = RetrievalMetrics.compute_metrics(v2t)                                    # as eval_epoch calls it (:250)
    assert vt["cols"] == gold[f"{name}_vt_cols"].tolist()
    np.testing.assert_array_equal(_vec(vt, VT_KEYS), gold[f"{name}_vt"])
    tv2, vt2 = multi_sentence_metrics(sim, cut)                                    # un-padded path, numpy input
    assert tv2 == tv and vt2["cols"] == vt["cols"]
    tv3, vt3 = multi_sentence_metrics(torch.from_numpy(sim).cuda(), cut.tolist())  # CUDA input
    assert tv3 == tv and vt3 == vt2


@pytest.mark.gpu
@pytest.mark.parametrize("V,maxlen,ties,nonfinite", [(670, 9, True, False), (672, 5, False, True), (1, 3, True, False),
                                                     (131, 1, True, True), (257, 40, True, False),
                                                     (4101, 2, True, True)])
def test_gpu_multi_sentence_kernels_vs_counting_oracle(V, maxlen, ties, nonfinite):
    """Kernels through ops (C ABI) against the numpy counting oracle at shapes covering the scalar (V % 4 != 0)
    and float4 paths, partial 128-column / 8-group tiles, one caption per video, one video, long caption groups,
    ties (stable column order) and NaN/inf scores, rows above 4096 columns (CTA-per-row variant; warp-per-row
    below); plus the column-sharded accumulate form."""
    from neighborretr_b200 import ops
    from neighborretr_b200.metrics import group_layout
    sim, cut = synth.make_multi_sentence_case(V, maxlen, ties, nonfinite, seed=100 + V)
    T = sim.shape[0]
    gs, tgt = group_layout(cut, total_rows=T)
    s = torch.from_numpy(sim).cuda()
    tg = torch.from_numpy(tgt).cuda()
    gt, eqb, valid = ops.rank_counts_target(s, tg)
    ranks, ok = OM.multi_sentence_ranks_by_counting(sim, cut)
    assert np.array_equal(valid.cpu().numpy().astype(bool), ok)
    got = (gt + eqb).cpu().numpy().astype(np.int64)
    assert np.array_equal(got[ok], ranks[ok])
    assert not got[~ok].any()                                             # invalid captions contribute nothing
    x = np.where(np.isnan(sim), -np.inf, sim)
    want = np.stack([x[gs[i]:gs[i + 1]].max(axis=0) for i in range(V)], axis=1)
    out = ops.group_max_t(s, torch.from_numpy(gs).cuda())
    assert tuple(out.shape) == (V, V) and np.array_equal(out.cpu().numpy(), want)
    if V >= 3:        # column shards: positives' scores passed in, counts accumulate across the shards
        diag = torch.from_numpy(sim[np.arange(T), tgt]).cuda()
        g2 = torch.zeros(T, dtype=torch.int32, device="cuda")
        e2 = torch.zeros(T, dtype=torch.int32, device="cuda")
        bounds = [0, V // 3, V // 3 + 1, V]
        for c0, c1 in zip(bounds[:-1], bounds[1:]):
            ops.rank_counts_target(s[:, c0:c1].contiguous(), tg, diag=diag, col_offset=c0, gt=g2, eq_before=e2,
                                   want_valid=False)
        assert torch.equal(g2, gt) and torch.equal(e2, eqb)


@pytest.mark.gpu
def test_gpu_rank_target_few_long_rows():
    """Few rows x many columns takes the CTA-per-row variant (Q < 4736 and N > 2048); unaligned row starts."""
    from neighborretr_b200 import ops
    rng = np.random.RandomState(3)
    Q, N = 300, 5001
    sim = rng.randint(0, 40, size=(Q, N)).astype(np.float32)
    sim[rng.rand(Q, N) < 0.002] = np.nan
    tgt = rng.randint(0, N, size=Q).astype(np.int32)
    sim[5, tgt[5]] = np.inf                                               # an invalid row
    gt, eqb, valid = ops.rank_counts_target(torch.from_numpy(sim).cuda(), torch.from_numpy(tgt).cuda())
    sd = sim[np.arange(Q), tgt][:, None]
    ok = np.isfinite(sd[:, 0])
    want_g = ((sim > sd) | np.isnan(sim)).sum(1)
    want_e = ((sim == sd) & (np.arange(N)[None, :] < tgt[:, None])).sum(1)
    assert np.array_equal(valid.cpu().numpy().astype(bool), ok) and not ok[5]
    assert np.array_equal(gt.cpu().numpy()[ok], want_g[ok]) and np.array_equal(eqb.cpu().numpy()[ok], want_e[ok])
    assert gt[5].item() == 0 and eqb[5].item() == 0


@pytest.mark.gpu
def test_gpu_multi_sentence_rejects_bad_shapes():
    from neighborretr_b200 import ops
    from neighborretr_b200.evaluator import multi_sentence_metrics
    from neighborretr_b200.metrics import RetrievalMetrics
    with pytest.raises(ValueError):
        RetrievalMetrics.tensor_text_to_video_metrics(np.zeros((3, 2, 4), np.float32))
    with pytest.raises(ValueError):
        multi_sentence_metrics(np.zeros((5, 3), np.float32), [1, 4])          # 2 cut-off points, 3 videos
    with pytest.raises(ValueError):
        ops.rank_counts_target(torch.zeros(4, 3, device="cuda"), torch.zeros(3, dtype=torch.int32, device="cuda"))
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        ops.group_max_t(torch.zeros(4, 3), torch.zeros(2, dtype=torch.int32))
