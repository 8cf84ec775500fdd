"""CPU-only checks: the C-ABI library builds, loads and exports every symbol include/nrhead.h declares (no compute
calls without a GPU); host-side logic (metrics expansion, synthetic inputs, no-CPU-fallback guards)."""
import os
import re

import numpy as np
import pytest
import torch

from neighborretr_b200 import _lib, synth
from oracle import metrics as OM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "nrhead.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libnrhead.so lacks {s}"
    assert set(syms) == set(_lib.SIGNATURES), (set(syms) ^ set(_lib.SIGNATURES))
    assert lib.nr_version() == 1
    assert lib.nr_prep_partials(3072) == 296 and lib.nr_prep_partials(5) == 1
    assert lib.nr_sinkhorn_workspace_bytes(128) >= 4


def test_bad_arguments_return_errors_without_a_gpu():
    lib = _lib.load()
    # argument validation happens before any CUDA call
    rc = lib.nr_maxsim_fwd(0, None, None, None, None, None, 0, 4, 4, 4, 64, 1.0, None, 0, 0, None, 0, 0, 0, None,
                           None, None)
    assert rc != 0 and b"empty problem" in lib.nr_last_error()
    rc = lib.nr_row_losses_fwd(1, 10, None, 0, None, None, None, None, 10, 10, 0, None, 20, 1.0, 1.0, 0.5, 2, 1, 1, 1,
                               None)
    assert rc != 0      # k=20 needs B >= 22 (reference crashes at until_module.py:123)
    assert b"cbank" in lib.nr_last_error() or b"num_neighbors" in lib.nr_last_error()


def test_ops_refuse_cpu_tensors():
    from neighborretr_b200 import ops, until_module as U
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        ops.Prepared(torch.randn(2, 3, 8))
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        U.CentralityWeightingLoss()(torch.randn(4, 4), torch.ones(4))


def test_metrics_from_counts_matches_sort_form():
    from neighborretr_b200.metrics import metrics_from_counts
    rng = np.random.RandomState(0)
    for mat in (rng.randn(50, 50).astype(np.float32), rng.randint(0, 4, (40, 40)).astype(np.float32)):
        g, e = OM.ranks_by_counting(mat)
        a, b = metrics_from_counts(g, e), OM.compute_metrics(mat)
        assert a["cols"] == b["cols"]
        for k in ("R1", "R5", "R10", "R50", "MR", "MedianR", "MeanR"):
            assert a[k] == b[k]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_counts_add_up_to_the_reference_ranks(world):
    """Column-sharded evaluation (evaluator.sharded_retrieval, fused or not): per-shard (greater, equal) counts against
    the all-reduced diagonal, summed over shards for text->video and concatenated for video->text, give exactly the
    `cols` of compute_metrics on the full matrix and on its transpose (reference utils/metrics.py:58-66)."""
    from neighborretr_b200.metrics import metrics_from_counts
    rng = np.random.RandomState(world)
    for mat in (rng.randn(37, 37).astype(np.float32), rng.randint(0, 3, (29, 29)).astype(np.float32)):
        n = mat.shape[0]
        per = (n + world - 1) // world
        gt_t = np.zeros(n, np.int64); eq_t = np.zeros(n, np.int64)
        gt_v = np.zeros(n, np.int64); eq_v = np.zeros(n, np.int64)
        for r in range(world):
            lo, hi = min(r * per, n), min((r + 1) * per, n)
            a, b, c, d = OM.shard_counts(mat, lo, hi)
            gt_t += a; eq_t += b
            gt_v[lo:hi], eq_v[lo:hi] = c, d
        assert metrics_from_counts(gt_t, eq_t)["cols"] == OM.compute_metrics(mat)["cols"]
        assert metrics_from_counts(gt_v, eq_v)["cols"] == OM.compute_metrics(mat.T)["cols"]


def test_synthetic_inputs_are_deterministic_and_ragged():
    a = synth.make_batch(16, 24, 12, d=64, seed=5, rank=1)
    b = synth.make_batch(16, 24, 12, d=64, seed=5, rank=1)
    assert torch.equal(a.text_feat, b.text_feat) and torch.equal(a.video_mask, b.video_mask)
    assert a.text_mask.dtype == torch.int64 and a.text_mask.sum(1).min() >= 3
    assert (a.text_mask[:, 0] == 1).all() and a.idx[0].item() == 16
    c = synth.make_batch(16, 24, 12, d=64, seed=5, rank=0)
    assert not torch.equal(a.text_feat, c.text_feat)


def test_reference_surface_names():
    """Same names / positional order as the reference modules."""
    import inspect
    from neighborretr_b200 import modeling as M, until_module as U, metrics as MT, evaluator as EV
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(M.HeadMixin.local_level) == ["self", "text_feat", "video_feat", "text_mask", "video_mask"]
    assert sig(M.HeadMixin.global_level) == ["self", "text_feat", "video_feat"]
    assert sig(M.HeadMixin._compute_losses)[:14] == [
        "self", "text_feat", "video_feat", "text_mask", "video_mask", "mb_feat_t", "mb_feat_v", "mb_mask_t",
        "mb_mask_v", "centrality_scale", "beta", "num_neighbors", "temperature", "logit_scale"]
    assert sig(M.HeadMixin.update_memory_bank) == ["self", "idx", "text_feat", "video_feat", "text_mask", "video_mask"]
    assert sig(M.HeadMixin.get_similarity_logits)[:5] == ["self", "text_feat", "video_feat", "text_mask", "video_mask"]
    assert sig(U.NeighborAdjustingLoss.forward) == ["self", "similarity_matrix", "memory_bank_matrix", "num_neighbors",
                                                    "temperature"]
    assert sig(U.UniformRegularizationLoss.forward) == ["self", "similarity_matrix", "logit_scale", "beta",
                                                        "num_iterations"]
    assert sig(U.CentralityWeightingLoss.forward) == ["self", "similarity_matrix", "centrality_weights"]
    assert sig(U.KLDivergenceLoss.forward) == ["self", "global_similarity", "local_similarity"]
    assert sig(EV._run_on_single_gpu) == ["model", "t_mask_list", "v_mask_list", "t_feat_list", "v_feat_list",
                                          "mini_batch"]
    assert sig(MT.RetrievalMetrics.compute_metrics) == ["similarity_matrix"]
    m = M.NeighborRetr(synth.default_config(), width=32)
    names = {n for n, _ in m.named_parameters()}
    for pre in ("text_weight_fc", "video_weight_fc", "text_weight_fc0", "video_weight_fc0", "text_weight_fc1",
                "video_weight_fc1", "text_weight_intra", "video_weight_intra"):
        for suf in ("0.weight", "0.bias", "2.weight", "2.bias"):
            assert f"{pre}.{suf}" in names
    assert "clip.logit_scale" in names


def test_ctypes_structs_match_the_c_header(tmp_path):
    """nr_maxsim2_problem / nr_maxsim2_bwd_job cross the C ABI by pointer: the ctypes mirrors in _lib.py must have
    the size and field offsets a C compiler gives the declarations of include/nrhead.h."""
    import ctypes
    import shutil
    import subprocess
    from neighborretr_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    structs = {"nr_maxsim2_problem": _lib.MaxSim2Problem, "nr_maxsim2_bwd_job": _lib.MaxSim2BwdJob,
               "nr_maxsim2_rank_problem": _lib.MaxSim2RankProblem, "nr_maxsim2_bwd_w_job": _lib.MaxSim2BwdWJob,
               "nr_softmax_side": _lib.SoftmaxSide, "nr_bank_side": _lib.BankSide, "nr_mlp_side": _lib.MlpSide}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nrhead.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call([gcc, "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).strip().splitlines()
    for line, (cname, cls) in zip(out, structs.items()):
        parts = line.split()
        assert parts[0] == cname
        assert int(parts[1]) == ctypes.sizeof(cls), cname
        assert [int(v) for v in parts[2:]] == [getattr(cls, f).offset for f, _ in cls._fields_], cname


def test_static_input_slab_views():
    """graph._slab_like: the seven step inputs as views of ONE byte slab (so that a staged batch moves in with a single
    copy): shapes / dtypes preserved, 256-byte aligned, non-overlapping, writes through the views land in the slab."""
    import torch
    from neighborretr_b200.graph import FIELDS, _slab_like
    h = synth.make_batch(5, 7, 3, d=16, seed=3)
    ex = [getattr(h, f) for f in FIELDS]
    slab, views = _slab_like(ex, torch.device("cpu"))
    assert slab.dtype == torch.uint8 and slab.dim() == 1
    spans = []
    for f, t in zip(FIELDS, ex):
        v = views[f]
        assert v.shape == t.shape and v.dtype == t.dtype and v.is_contiguous()
        off = v.data_ptr() - slab.data_ptr()
        assert off % 256 == 0
        spans.append((off, off + t.numel() * t.element_size()))
        v.copy_(t)
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= slab.numel()
    slab2, views2 = _slab_like(ex, torch.device("cpu"))
    slab2.copy_(slab)                          # a second slab filled by ONE copy of the first reproduces every input
    for f, t in zip(FIELDS, ex):
        assert torch.equal(views2[f], t)


def _band_decode(local, n_me, n_nt, wn):
    """Python restatement of the tile order of csrc/maxsim2_tc.cu (decode lambda): bands of wn Y tiles, inside a band
    the Y tile runs fastest."""
    band_tiles = n_me * wn
    bnd, rem = divmod(local, band_tiles)
    w = min(wn, n_nt - bnd * wn)
    mt = rem // w
    return mt, bnd * wn + rem - mt * w


def _diag_tile(local, SX, SY, Rx, Ry, gx0, gy0, dj):
    """... of diag_tile(): entry `local` = (X box, j-th Y tile holding positives of that box), or None."""
    mt, j = divmod(local, dj)
    off = gx0 - gy0
    x0 = mt * SX
    lo = max(x0 + off, 0)
    hi = min(min(x0 + SX, Rx) - 1 + off, Ry - 1)
    nt = lo // SY + j
    return (mt, nt) if hi >= lo and nt <= hi // SY else None


@pytest.mark.parametrize("n_me,n_nt,wn", [(1, 1, 1), (26, 26, 68), (1639, 410, 65), (7, 10, 3), (5, 9, 9), (3, 17, 16)])
def test_band_order_visits_every_tile_once(n_me, n_nt, wn):
    wn = min(wn, n_nt)
    seen = [_band_decode(t, n_me, n_nt, wn) for t in range(n_me * n_nt)]
    assert len(set(seen)) == n_me * n_nt
    assert all(0 <= mt < n_me and 0 <= nt < n_nt for mt, nt in seen)
    if wn == n_nt:                                  # a single band is the plain row-major order
        assert seen == [(t // n_nt, t % n_nt) for t in range(n_me * n_nt)]


@pytest.mark.parametrize("SX,SY,Rx,Ry,gx0,gy0", [(5, 20, 101, 101, 0, 0), (5, 20, 101, 37, 0, 40), (5, 20, 101, 6, 0, 95),
                                               (2, 4, 37, 37, 0, 0), (10, 10, 45, 45, 0, 0), (10, 10, 12, 45, 20, 0),
                                               (32, 5, 70, 70, 0, 0)])
def test_diag_tile_list_covers_every_positive_once(SX, SY, Rx, Ry, gx0, gy0):
    """Pass 1 of nr_maxsim2_rank contracts only tiles that hold a positive pair (pair id gx0+rx == gy0+ry): the tile
    list must contain every such tile exactly once and nothing else."""
    dj = (SX + SY - 2) // SY + 1
    n_mt = (Rx + SX - 1) // SX
    tiles = [t for t in (_diag_tile(l, SX, SY, Rx, Ry, gx0, gy0, dj) for l in range(n_mt * dj)) if t is not None]
    want = {(rx // SX, (rx + gx0 - gy0) // SY) for rx in range(Rx) if 0 <= rx + gx0 - gy0 < Ry}
    assert len(tiles) == len(set(tiles)) and set(tiles) == want
