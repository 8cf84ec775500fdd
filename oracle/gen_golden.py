"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.npz`` from the REFERENCE itself.

Run in the build container only (``python -m oracle.gen_golden``): imports the unmodified
reference from ``/root/reference`` (read-only; stubs for the absent ``timm``/``ftfy``/``boto3``
imports), builds a weights-free ``NeighborRetr`` shell, executes the
reference's own head functions on seeded synthetic inputs and stores their outputs.  The GPU box
has no ``/root/reference``; tests only read the committed ``.npz`` files.

The only intervention on the reference in the head cases is replacing ``merge_global_features``
(token clustering with ``torch.rand`` noise, cluster.py:483-484) by a function
returning the seeded global features, so that ``_compute_losses`` is deterministic.  The token
clustering itself is pinned separately (``run_cluster``: the reference's own CTM / TCBlock classes
under ``torch.manual_seed``), as are the multi-sentence metrics (``run_multi_sentence``) and the
memory-bank manager (``run_prefill``).
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from neighborretr_b200 import synth  # noqa: E402  (input plumbing only)

REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")


def import_reference():
    def stub(name, **attrs):
        if name in sys.modules:
            return
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m

    stub("timm")
    stub("timm.models")
    stub("timm.models.layers", drop_path=lambda x, p=0.0, t=False: x)
    stub("ftfy", fix_text=lambda s: s)
    stub("boto3")
    stub("botocore")
    stub("botocore.exceptions", ClientError=Exception)
    sys.path.insert(0, REF)
    from NeighborRetr.models import modeling, until_module
    from NeighborRetr.training import evaluator
    from NeighborRetr.utils import metrics
    return modeling, until_module, evaluator, metrics


def reference_head(modeling, d, cfg, params, logit_scale=float(np.log(100.0))):
    m = modeling.NeighborRetr.__new__(modeling.NeighborRetr)
    nn.Module.__init__(m)
    m.config = cfg
    m.transformer_width = d
    m._init_weighting_networks()
    m._init_loss_functions()
    m._init_memory_bank()
    m.apply(m._init_weights)
    for name, sd in params.items():
        getattr(m, name).load_state_dict(sd)
    holder = nn.Module()
    holder.logit_scale = nn.Parameter(torch.tensor(logit_scale))
    m.clip = holder
    return m


CASES = {
    # name: dict(b, nt, nv, d, m, k, ragged_bank)
    "small": dict(b=40, nt=8, nv=6, d=64, m=56, k=20),
    "small_k8": dict(b=48, nt=10, nv=5, d=32, m=32, k=8),
    "cfg1": dict(b=128, nt=24, nv=12, d=512, m=512, k=20),
}


def make_case(c, dtype=torch.float32):
    h = synth.make_batch(c["b"], c["nt"], c["nv"], d=c["d"], seed=1234)
    bank = synth.make_bank(c["m"], c["nt"], c["nv"], d=c["d"])
    params = synth.make_mlp_params(d=c["d"])
    cfg = synth.default_config(num_neighbors=c["k"])
    return h, bank, params, cfg


def sub(t, n=4096):
    """Strided subsample for large gradient tensors (deterministic)."""
    f = t.reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].clone()


def run_losses(modeling, c, name):
    h, bank, params, cfg = make_case(c)
    m = reference_head(modeling, c["d"], cfg, params)
    text = h.text_feat.clone().requires_grad_(True)
    video = h.video_feat.clone().requires_grad_(True)
    gt = h.global_text.clone().requires_grad_(True)
    gv = h.global_video.clone().requires_grad_(True)
    m.merge_global_features = lambda *a, **k: (gt, gv)
    out = {}
    # piecewise outputs
    s, _ = m.local_level(text, video, h.text_mask, h.video_mask)
    g, _ = m.global_level(gt, gv)
    wt, wv = m.compute_centrality_weights(text, video, gt, gv, cfg.centrality_scale)
    mb_t2v, _ = m.local_level(text, bank.mb_feat_v, h.text_mask, bank.mb_mask_v)
    _, mb_v2t = m.local_level(bank.mb_feat_t, video, bank.mb_mask_t, h.video_mask)
    out.update(S=s, G=g, w_t=wt, w_v=wv, mb_t2v=mb_t2v, mb_v2t=mb_v2t)
    ls = m.clip.logit_scale.exp()
    out["Lc_t2v"] = m.centrality_weighting_loss(s * ls, wt)
    out["Lc_v2t"] = m.centrality_weighting_loss(s.T * ls, wv)
    out["Ln_t2v"] = m.neighbor_adjusting_loss(s, mb_v2t, cfg.num_neighbors, cfg.temperature)
    out["Ln_v2t"] = m.neighbor_adjusting_loss(s.T, mb_t2v, cfg.num_neighbors, cfg.temperature)
    out["Lu_t2v"] = m.uniform_regularization_loss(g, cfg.temperature, cfg.beta)
    out["Lu_v2t"] = m.uniform_regularization_loss(g.T, cfg.temperature, cfg.beta)
    out["Lkl_t2v"] = m.kl_loss(g, s)
    out["Lkl_v2t"] = m.kl_loss(g.T, s.T)
    out["sinkhorn_T"] = m.uniform_regularization_loss.sinkhorn_algorithm(g, cfg.beta, 50)
    nb, ext = m.neighbor_adjusting_loss.create_neighbor_mask(s, cfg.num_neighbors)
    out["nbr_mask"] = nb.to(torch.uint8)
    # full head fwd + bwd
    losses = m._compute_losses(text, video, h.text_mask, h.video_mask,
                               bank.mb_feat_t, bank.mb_feat_v, bank.mb_mask_t, bank.mb_mask_v,
                               cfg.centrality_scale, cfg.beta, cfg.num_neighbors, cfg.temperature,
                               m.clip.logit_scale.exp())
    out["losses"] = torch.stack([x.detach() for x in losses])
    losses[0].backward()
    big = text.numel() > 200_000
    pick = sub if big else (lambda t: t.clone())
    out["subsampled"] = torch.tensor(int(big))
    out["g_text"] = pick(text.grad)
    out["g_video"] = pick(video.grad)
    out["g_gt"] = pick(gt.grad)
    out["g_gv"] = pick(gv.grad)
    out["g_logit_scale"] = m.clip.logit_scale.grad
    out["gn_text"] = text.grad.norm()
    out["gn_video"] = video.grad.norm()
    for nme in ("text_weight_fc", "video_weight_fc"):
        for pn, p in getattr(m, nme).named_parameters():
            out[f"g_{nme}.{pn}"] = pick(p.grad)
            out[f"gn_{nme}.{pn}"] = p.grad.norm()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"),
                        **{k: v.detach().cpu().numpy() for k, v in out.items()})
    print(name, "losses", out["losses"].tolist())


def run_act_piece(modeling):
    """ActivityNet-shaped a1 only (the reference head crashes at these shapes)."""
    c = dict(b=24, nt=64, nv=64, d=512, m=8, k=20)
    h, bank, params, cfg = make_case(c)
    m = reference_head(modeling, c["d"], cfg, params)
    with torch.no_grad():
        s, _ = m.local_level(h.text_feat, h.video_feat, h.text_mask, h.video_mask)
    np.savez_compressed(os.path.join(OUT, "act_piece.npz"), S=s.numpy())
    print("act_piece", s.shape)


def run_eval(modeling, evaluator, metrics):
    c = dict(b=100, nt=8, nv=6, d=64, m=8, k=20)
    h, bank, params, cfg = make_case(c)
    m = reference_head(modeling, c["d"], cfg, params)
    m.eval()
    nv = 70
    sim, sim_t = evaluator._run_on_single_gpu(m, h.text_mask, h.video_mask[:nv],
                                              h.text_feat, h.video_feat[:nv], mini_batch=64)
    out = {"sim_100x70": sim}
    # metrics: random square, and an integer-valued matrix with many ties
    rng = np.random.RandomState(7)
    cases = {
        "rand": rng.randn(64, 64).astype(np.float32),
        "ties": rng.randint(0, 6, size=(48, 48)).astype(np.float32),
        "sim": evaluator._run_on_single_gpu(m, h.text_mask, h.video_mask, h.text_feat,
                                            h.video_feat, mini_batch=64)[0],
    }
    for k, mat in cases.items():
        r = metrics.RetrievalMetrics.compute_metrics(mat)
        out[f"{k}_mat"] = mat
        out[f"{k}_cols"] = np.asarray(r["cols"], dtype=np.int64)
        out[f"{k}_scalars"] = np.asarray([r["R1"], r["R5"], r["R10"], r["R50"], r["MR"],
                                          r["MedianR"], r["MeanR"]], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "eval.npz"), **out)
    print("eval ok")


def run_multi_sentence(metrics):
    """Multi-sentence metrics of the reference (utils/metrics.py:81-145) on the padded tensor built the way
    eval_epoch builds it (training/evaluator.py:216-239; that code is inline in eval_epoch, so the padding is
    restated in oracle.metrics.multi_sentence_reshape and only the metric functions are the reference's)."""
    from oracle import metrics as OM
    R = metrics.RetrievalMetrics
    out = {}
    for name, spec in synth.MS_CASES.items():
        sim, cut = synth.make_multi_sentence_case(*spec)
        pad = OM.multi_sentence_reshape(sim, cut)
        tv = R.tensor_text_to_video_metrics(pad.copy())
        v2t = R.tensor_video_to_text_sim(torch.tensor(pad.copy())).numpy()
        vt = R.compute_metrics(v2t)
        keys = ["R1", "R5", "R10", "R50", "MedianR", "MeanR", "Std_Rank", "MR"]
        out[f"{name}_tv"] = np.asarray([tv[k] for k in keys], dtype=np.float64)
        out[f"{name}_v2t_sim"] = v2t
        out[f"{name}_vt_cols"] = np.asarray(vt["cols"], dtype=np.int64)
        out[f"{name}_vt"] = np.asarray([vt[k] for k in ("R1", "R5", "R10", "R50", "MR", "MedianR", "MeanR")],
                                       dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "multi_sentence.npz"), **out)
    print("multi_sentence ok")


def run_prefill():
    """MemoryBankManager.load_memory_bank / clear_memory_bank of the reference (utils/memory_bank.py) on a toy
    encoder and a list-of-batches loader (CPU, single process): 6 loader batches, mb_batch = 4."""
    import logging
    from NeighborRetr.utils.memory_bank import MemoryBankManager
    args = SimpleNamespace(logger=logging.getLogger("golden"), mb_batch=4, batch_size=5, distributed=False,
                           world_size=1, local_rank=0)
    model = synth.ToyEncoder(d=8)
    n = MemoryBankManager(args).load_memory_bank(model, synth.make_prefill_loader(6, 5), "cpu", 0)
    out = {"rows": np.asarray(n), "ind": model.mb_ind.numpy(), "feat_t": model.mb_feat_t.numpy(),
           "feat_v": model.mb_feat_v.numpy(), "mask_t": model.mb_mask_t.numpy(), "mask_v": model.mb_mask_v.numpy(),
           "mb_batch": np.asarray(model.mb_batch)}
    MemoryBankManager(args).clear_memory_bank(model)
    out["cleared_shapes"] = np.asarray([model.mb_ind.numel(), model.mb_feat_t.dim(), model.mb_mask_v.dim(),
                                        model.mb_batch])
    np.savez_compressed(os.path.join(OUT, "prefill.npz"), **out)
    print("prefill ok", n)


CLUSTER_CASES = {
    # name: (batch, words, frames, width, heads)
    "msr": (12, 24, 12, 32, 8),        # merges to one global token per modality
    "act": (6, 64, 64, 32, 4),         # 64 words / 64 frames: 3 text and 6 video global tokens
}


def run_cluster(modeling):
    """Token clustering of the reference: ``NeighborRetr.merge_global_features`` (modeling.py:446-481) on a shell
    whose eight CTM / TCBlock layers are the reference's own classes (cluster.py) at a small width.  Stores the
    layers' state_dict, the outputs and gradients; the torch.rand tie-break noise is pinned by ``manual_seed``."""
    from NeighborRetr.models import cluster as C
    out = {}
    for name, (b, nt, nv, d, heads) in CLUSTER_CASES.items():
        torch.manual_seed(7)
        shell = nn.Module()
        for mod, (r0, r1) in (("text", (1 / 6, 1 / 4)), ("video", (1 / 4, 1 / 3))):     # modeling.py:186-197
            setattr(shell, f"{mod}_ctm0", C.CTM(sample_ratio=r0, embed_dim=d, dim_out=d, k=3))
            setattr(shell, f"{mod}_block0", C.TCBlock(dim=d, num_heads=heads))
            setattr(shell, f"{mod}_ctm1", C.CTM(sample_ratio=r1, embed_dim=d, dim_out=d, k=3))
            setattr(shell, f"{mod}_block1", C.TCBlock(dim=d, num_heads=heads))
        for p in shell.parameters():              # non-trivial biases / LayerNorm gains
            if p.dim() == 1:
                p.data.add_(0.1 * torch.randn_like(p))
        h = synth.make_batch(b, nt, nv, d=d, seed=2024)
        text = (h.text_feat / 6).clone().requires_grad_(True)
        video = (h.video_feat / 6).clone().requires_grad_(True)
        torch.manual_seed(123)
        gt, gv = modeling.NeighborRetr.merge_global_features(shell, text, video, h.text_mask, h.video_mask)
        wt = torch.linspace(-1, 1, gt.numel()).view_as(gt)
        wv = torch.linspace(1, -1, gv.numel()).view_as(gv)
        ((gt * wt).sum() + (gv * wv).sum()).backward()
        out[f"{name}_gt"], out[f"{name}_gv"] = gt.detach().numpy(), gv.detach().numpy()
        out[f"{name}_dtext"], out[f"{name}_dvideo"] = text.grad.numpy(), video.grad.numpy()
        for k, v in shell.state_dict().items():
            out[f"{name}_p_{k}"] = v.numpy()
        for k, p in shell.named_parameters():
            out[f"{name}_g_{k}"] = p.grad.numpy() if p.grad is not None else np.zeros(0, np.float32)
        print("cluster", name, tuple(gt.shape), tuple(gv.shape))
    np.savez_compressed(os.path.join(OUT, "cluster.npz"), **out)


def run_bank(modeling):
    c = dict(b=6, nt=4, nv=3, d=8, m=8, k=20)
    cfg = synth.default_config()
    m = reference_head(modeling, c["d"], cfg, synth.make_mlp_params(d=c["d"]))
    out = {}
    for step in range(4):
        h = synth.make_batch(c["b"], c["nt"], c["nv"], d=c["d"], seed=50 + step, rank=step)
        if step == 1:   # externally assigned prefill (memory_bank.py:206-211): capacity 14
            bk = synth.make_bank(14, c["nt"], c["nv"], d=c["d"])
            m.mb_ind, m.mb_feat_t, m.mb_feat_v = bk.mb_ind, bk.mb_feat_t, bk.mb_feat_v
            m.mb_mask_t, m.mb_mask_v = bk.mb_mask_t, bk.mb_mask_v
        m.update_memory_bank(h.idx, h.text_feat, h.video_feat, h.text_mask, h.video_mask)
        out[f"ind_{step}"] = m.mb_ind.numpy().copy()
        out[f"feat_v_{step}"] = m.mb_feat_v.numpy().copy()
        out[f"mask_t_{step}"] = m.mb_mask_t.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "bank.npz"), **out)
    print("bank ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    modeling, until_module, evaluator, metrics = import_reference()
    for name, c in CASES.items():
        run_losses(modeling, c, name)
    run_act_piece(modeling)
    run_eval(modeling, evaluator, metrics)
    run_bank(modeling)
    run_multi_sentence(metrics)
    run_prefill()
    run_cluster(modeling)


if __name__ == "__main__":
    main()
