"""B200-native retrieval head for NeighborRetr.

Public surface mirrors the reference's module names:
    neighborretr_b200.modeling.NeighborRetr / HeadMixin
    neighborretr_b200.until_module.{CentralityWeightingLoss, NeighborAdjustingLoss,
                                    UniformRegularizationLoss, KLDivergenceLoss, AllGather, AllGather2}
    neighborretr_b200.metrics.RetrievalMetrics.{compute_metrics, tensor_text_to_video_metrics,
                                                tensor_video_to_text_sim}
    neighborretr_b200.evaluator.{_run_on_single_gpu, multi_sentence_metrics, gather_eval_features}
    neighborretr_b200.memory_bank.MemoryBankManager
    neighborretr_b200.install(...)  -> rebind the above onto an imported reference checkout
"""
__all__ = ["install", "bind_head"]


def bind_head(cls, graph=True, precision=None):
    """Rebind the head onto ONE class with the reference model's shape (``get_text_video_feat``,
    ``merge_global_features``, ``config``, ``clip.logit_scale``, the two token-weight MLPs and the mb_* attributes —
    reference modeling.py:137-197, :541-585): every HeadMixin method plus ``forward``.  install() calls this on the
    reference's NeighborRetr; tests call it on a from-scratch stand-in.  Returns the rebound method names."""
    from . import modeling as md
    names = []
    for name, fn in vars(md.HeadMixin).items():
        if (callable(fn) or isinstance(fn, property)) and not name.startswith("__"):
            setattr(cls, name, fn)              # methods, and the five mb_* bank properties
            names.append(name)
    cls.forward = md.installed_forward
    cls.head_graph = bool(graph)
    cls.head_precision = precision or md.INSTALL_PRECISION
    return names + ["forward"]


def install(reference_pkg=None, graph=True, precision=None):
    """Rebind the CUDA head onto the already-importable reference package ``NeighborRetr`` so that the
    reference's main.py / training loop run unchanged.  Returns the list of patched names.

    ``forward`` itself is replaced (reference modeling.py:251-312): the reference's encoders, then
    ``head_forward`` — the row-block sharded head at world_size > 1 instead of five gathers + a replicated head.
    graph=True: every training step of the head (forward + backward + bank FIFO) is one CUDA-graph replay whose
    gradients are handed to autograd by ``loss.backward()`` (graph.GraphedHead).
    precision: arithmetic of the token-pair contraction for the installed class, "bf16x3" (default: split-bf16
    tensor-core products, fp32-accurate), "bf16" (fastest, 1e-2 loss tolerance) or "fp32" (CUDA cores)."""
    import importlib

    from . import evaluator as ev
    from . import metrics as mt
    from . import modeling as md
    from . import until_module as um

    ref_modeling = importlib.import_module("NeighborRetr.models.modeling")
    ref_until = importlib.import_module("NeighborRetr.models.until_module")
    ref_metrics = importlib.import_module("NeighborRetr.utils.metrics")
    ref_eval = importlib.import_module("NeighborRetr.training.evaluator")
    patched = []
    cls = ref_modeling.NeighborRetr
    # every head method (public ones with the reference's names, plus the private helpers they call) and forward
    patched += [f"NeighborRetr.models.modeling.NeighborRetr.{n}" for n in bind_head(cls, graph, precision)]
    for name in ("CentralityWeightingLoss", "NeighborAdjustingLoss", "UniformRegularizationLoss",
                 "KLDivergenceLoss", "AllGather", "AllGather2"):
        for mod in (ref_until, ref_modeling):
            setattr(mod, name, getattr(um, name))
        patched.append(f"NeighborRetr.models.until_module.{name}")
    ref_modeling.allgather = um.AllGather.apply
    ref_modeling.allgather2 = um.AllGather2.apply
    ref_eval.AllGather = um.AllGather
    ref_eval.allgather = um.AllGather.apply
    for name in ("compute_metrics", "tensor_text_to_video_metrics", "tensor_video_to_text_sim"):
        setattr(ref_metrics.RetrievalMetrics, name, staticmethod(getattr(mt.RetrievalMetrics, name)))
        patched.append(f"NeighborRetr.utils.metrics.RetrievalMetrics.{name}")
    ref_eval._run_on_single_gpu = ev._run_on_single_gpu
    patched += ["NeighborRetr.training.evaluator._run_on_single_gpu"]
    # methods are rebound on the reference's class (main.py imports the class by name before this runs)
    from . import memory_bank as mb
    ref_mb = importlib.import_module("NeighborRetr.utils.memory_bank")
    for name in ("load_memory_bank", "clear_memory_bank", "_info", "_error"):
        setattr(ref_mb.MemoryBankManager, name, getattr(mb.MemoryBankManager, name))
    patched += ["NeighborRetr.utils.memory_bank.MemoryBankManager.load_memory_bank",
                "NeighborRetr.utils.memory_bank.MemoryBankManager.clear_memory_bank"]
    return patched
