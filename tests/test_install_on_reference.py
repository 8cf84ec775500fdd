"""neighborretr_b200.install() against the real reference checkout (only where /root/reference exists, i.e. in the
build container; skipped on the GPU box).  CPU-only: checks that every documented name is rebound on the imported
reference package, that the rebound head refuses CPU tensors (no fallback), and that the pieces that are pure
host data movement (memory-bank prefill) run through the reference's own class after the rebinding."""
import logging
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import load_golden
from neighborretr_b200 import synth

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "NeighborRetr")),
                                reason="reference checkout not present")


@pytest.fixture(scope="module")
def installed():
    from oracle.gen_golden import import_reference, reference_head
    modeling, until_module, evaluator, metrics = import_reference()
    import neighborretr_b200
    patched = neighborretr_b200.install()
    return SimpleNamespace(modeling=modeling, until_module=until_module, evaluator=evaluator, metrics=metrics,
                           patched=patched, reference_head=reference_head)


def test_install_rebinds_the_documented_names(installed):
    from neighborretr_b200 import evaluator as EV, memory_bank as MB, metrics as MT, modeling as MD, until_module as UM
    ref = installed
    for name in ("local_level", "global_level", "get_similarity_logits", "compute_centrality_weights",
                 "_compute_losses", "update_memory_bank"):
        assert getattr(ref.modeling.NeighborRetr, name) is getattr(MD.HeadMixin, name), name
    for name in ("CentralityWeightingLoss", "NeighborAdjustingLoss", "UniformRegularizationLoss", "KLDivergenceLoss",
                 "AllGather", "AllGather2"):
        assert getattr(ref.until_module, name) is getattr(UM, name)
        assert getattr(ref.modeling, name) is getattr(UM, name)
    assert ref.evaluator._run_on_single_gpu is EV._run_on_single_gpu
    R = ref.metrics.RetrievalMetrics
    for name in ("compute_metrics", "tensor_text_to_video_metrics", "tensor_video_to_text_sim"):
        assert getattr(R, name) is getattr(MT.RetrievalMetrics, name)
        assert f"NeighborRetr.utils.metrics.RetrievalMetrics.{name}" in ref.patched
    import NeighborRetr.utils.memory_bank as ref_mb
    assert ref_mb.MemoryBankManager.load_memory_bank is MB.MemoryBankManager.load_memory_bank
    # the reference's merge_global_features (token clustering) and encoders stay the reference's own
    assert ref.modeling.NeighborRetr.merge_global_features.__module__ == "NeighborRetr.models.modeling"


def test_installed_head_has_no_cpu_fallback(installed):
    c = dict(b=24, nt=6, nv=4, d=32, m=8, k=20)
    cfg = synth.default_config()
    head = installed.reference_head(installed.modeling, c["d"], cfg, synth.make_mlp_params(d=c["d"]))
    h = synth.make_batch(c["b"], c["nt"], c["nv"], d=c["d"], seed=3)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA tensors required|no CPU fallback"):
            head.local_level(h.text_feat, h.video_feat, h.text_mask, h.video_mask)
        with pytest.raises(RuntimeError, match="CUDA tensors required|no CPU fallback"):
            installed.metrics.RetrievalMetrics.compute_metrics(np.eye(4, dtype=np.float32))


def test_installed_memory_bank_manager_runs_through_the_reference_class(installed):
    import NeighborRetr.utils.memory_bank as ref_mb
    gold = load_golden("prefill")
    args = SimpleNamespace(logger=logging.getLogger("test"), mb_batch=4, batch_size=5, distributed=False, world_size=1,
                           local_rank=0)
    mgr = ref_mb.MemoryBankManager(args)                 # the reference's constructor, our load/clear
    model = synth.ToyEncoder(d=8)
    assert mgr.load_memory_bank(model, synth.make_prefill_loader(6, 5), "cpu", 0) == 20
    assert np.array_equal(model.mb_feat_t.numpy(), gold["feat_t"]) and np.array_equal(model.mb_ind.numpy(), gold["ind"])
    mgr.clear_memory_bank(model)
    assert model.mb_batch == 0
