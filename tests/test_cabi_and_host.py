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
               "nr_maxsim2_rank_problem": _lib.MaxSim2RankProblem, "nr_maxsim2_bwd_w_job": _lib.MaxSim2BwdWJob}
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
