"""Parity of the multi-rank (row-block sharded, CUDA-graph captured) head step against the single-process
full-batch head on the rank-ordered concatenation of every rank's batch — what the reference evaluates on every
rank after its gathers (reference NeighborRetr/models/modeling.py:274-312, until_module.py:367-388).

Used by ``bench.py --check`` (on by default at N > 1: the step that is TIMED is the step that is checked) and by
tests/dist_graph_check.py.  Both sides run on the CUDA path; absolute parity of the single-process head against the
reference is pinned separately (tests/test_gpu_parity.py, tests/test_zz_fullsize.py).

Checked per rank: the five losses, this rank's rows of the feature / global-feature gradients, every head-parameter
gradient and the logit_scale gradient (full gradients on every rank, as in the reference's replicated head), and the
memory bank after the step (FIFO of the gathered batch).
"""
from __future__ import annotations

import torch

from . import synth

PARAM_NAMES = ("text_weight_fc", "video_weight_fc")


def _rel_l2(a, b):
    a, b = a.detach().double().reshape(-1), b.detach().double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def make_model(cfg, dev, precision, bwd_precision=None, d=512, mlp_precision=None):
    from .modeling import NeighborRetr
    model = NeighborRetr(cfg, width=d)
    for name, sd in synth.make_mlp_params(d=d).items():
        getattr(model, name).load_state_dict(sd)
    model.clip.logit_scale.data.fill_(float(torch.log(torch.tensor(100.0))))
    model.head_precision = precision
    model.head_bwd_precision = bwd_precision or precision
    if mlp_precision:
        model.head_mlp_precision = mlp_precision
    return model.to(dev).train()


def set_bank(model, bank, dev):
    for n in ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v"):
        setattr(model, n, getattr(bank, n).to(dev).clone())
    model.mb_batch = bank.mb_ind.shape[0]


def full_batch_reference(dev, world, shape, b, mrows, precision, bwd_precision=None, d=512, seed=1234,
                         mlp_precision=None):
    """Single-process head (world_size 1) on the concatenation of all ranks' synthetic batches (bench.py draws rank
    r's batch from seed 1234 + r).  Returns a dict of CUDA tensors."""
    nt, nv, _ = synth.SHAPES[shape]
    parts = [synth.make_batch(b, nt, nv, d=d, seed=seed, rank=r) for r in range(world)]
    cat = lambda f: torch.cat([getattr(p, f) for p in parts]).to(dev)
    model = make_model(synth.default_config(), dev, precision, bwd_precision, d, mlp_precision)
    set_bank(model, synth.make_bank(mrows, nt, nv, d=d), dev)
    text = cat("text_feat").requires_grad_(True)
    video = cat("video_feat").requires_grad_(True)
    gt = cat("global_text").requires_grad_(True)
    gv = cat("global_video").requires_grad_(True)
    losses = model.head_forward(text, video, cat("text_mask"), cat("video_mask"), cat("idx"), global_feats=(gt, gv))
    losses[0].backward()
    out = {"losses": torch.stack([x.detach() for x in losses]), "text": text.grad, "video": video.grad,
           "global_text": gt.grad, "global_video": gv.grad, "logit_scale": model.clip.logit_scale.grad}
    for n in PARAM_NAMES:
        for pn, p in getattr(model, n).named_parameters():
            out[f"{n}.{pn}"] = p.grad
    for n in ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v"):
        out[n] = getattr(model, n)
    torch.cuda.synchronize(dev)
    return out


def compare_step(ref, rank, b, losses, grads, model):
    """ref: full_batch_reference(); losses [5]; grads: dict text_feat/video_feat/global_text/global_video -> this
    rank's gradient rows; model: the sharded model after ONE step from the same initial bank.  Returns the error
    dict (max relative loss error, per-tensor rel-L2, bank equality)."""
    sl = slice(rank * b, (rank + 1) * b)
    e = {"loss_rel": float(((losses.detach() - ref["losses"]).abs() / ref["losses"].abs().clamp_min(1e-12)).max())}
    g = {"text": _rel_l2(grads["text_feat"], ref["text"][sl]), "video": _rel_l2(grads["video_feat"], ref["video"][sl]),
         "global_text": _rel_l2(grads["global_text"].reshape(b, -1), ref["global_text"][sl].reshape(b, -1)),
         "global_video": _rel_l2(grads["global_video"].reshape(b, -1), ref["global_video"][sl].reshape(b, -1)),
         "logit_scale": _rel_l2(model.clip.logit_scale.grad, ref["logit_scale"])}
    for n in PARAM_NAMES:
        for pn, p in getattr(model, n).named_parameters():
            if pn == "2.bias":
                # the softmax over tokens is invariant to the second layer's bias: its gradient is 0 up to rounding,
                # so it is compared on the scale of the second layer's weight gradient instead of relatively
                scale = ref[f"{n}.2.weight"].double().norm().clamp_min(1e-30)
                g[f"{n}.{pn}"] = float((p.grad.double() - ref[f"{n}.{pn}"].double()).norm() / scale)
            else:
                g[f"{n}.{pn}"] = _rel_l2(p.grad, ref[f"{n}.{pn}"])
    e["grads"] = g
    e["grad_rel_l2"] = max(g.values())
    e["bank_equal"] = all(bool(torch.equal(getattr(model, n), ref[n]))
                          for n in ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v"))
    return e


def tolerances(precision):
    """(loss relative, gradient rel-L2).  Both sides use the same kernels; they differ in tile composition, split-K
    and atomic accumulation order, so bf16 runs agree to the rounding of the bf16 routing coefficients."""
    if precision == "fp32":
        return 1e-5, 2e-4
    if precision == "bf16x3":
        return 1e-5, 2e-3
    return 1e-3, 2e-2


def check_all_ranks(err, precision, dev):
    """All-reduce the verdict and the worst errors over ranks; returns (ok, summary dict for the bench line)."""
    import torch.distributed as dist
    ltol, gtol = tolerances(precision)
    ok = err["loss_rel"] < ltol and err["grad_rel_l2"] < gtol and err["bank_equal"]
    t = torch.tensor([err["loss_rel"], err["grad_rel_l2"], 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    worst = max(err["grads"], key=err["grads"].get)
    return t[2].item() == 0.0, {"world": world, "loss_rel": t[0].item(), "grad_rel_l2": t[1].item(),
                                "bank_equal": bool(err["bank_equal"]), "loss_tol": ltol, "grad_tol": gtol,
                                "worst_tensor_rank0": worst,
                                "against": "single-process full-batch head on the concatenated batch (same kernels, "
                                           "world_size 1)"}
