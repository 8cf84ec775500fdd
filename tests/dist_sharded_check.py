"""Multi-rank check of the row-block sharded head (run under torchrun, or spawned by test_gpu_sharded.py):
every rank's losses equal the single-process full-batch head; its feature / global-feature gradients equal its rows
of the full-batch gradient; head-parameter and logit_scale gradients equal the full gradients on every rank; the
memory bank after the step equals the reference FIFO on the gathered batch.

    torchrun --nproc-per-node 2 tests/dist_sharded_check.py [--backend nccl|gloo] [--precision fp32|bf16]
With --backend gloo all ranks may share GPU 0 (single-GPU emulation; collectives are staged through the host)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--b", type=int, default=32)
    ap.add_argument("--shape", default="msrvtt")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    ndev = torch.cuda.device_count()
    dev = torch.device("cuda", local % ndev)
    torch.cuda.set_device(dev)
    dist.init_process_group(a.backend, rank=rank, world_size=world)
    from helpers import make_head, rel_l2, set_bank
    from neighborretr_b200 import synth
    nt, nv, _ = synth.SHAPES[a.shape]
    d, M, b = 512, 96, a.b
    params = synth.make_mlp_params(d=d)
    bank = synth.make_bank(M, nt, nv, d=d)
    parts = [synth.make_batch(b, nt, nv, d=d, seed=1234, rank=r) for r in range(world)]
    cat = lambda f: torch.cat([getattr(p, f) for p in parts]).to(dev)

    def run(model, h_text, h_video, tm, vm, idx, gt, gv):
        text = h_text.clone().requires_grad_(True); video = h_video.clone().requires_grad_(True)
        gtt = gt.clone().requires_grad_(True); gvv = gv.clone().requires_grad_(True)
        model.zero_grad(set_to_none=True)
        losses = model.head_forward(text, video, tm, vm, idx, global_feats=(gtt, gvv))
        losses[0].backward()
        return torch.stack([x.detach() for x in losses]), text.grad, video.grad, gtt.grad, gvv.grad

    # single-process full-batch head (world_size=1 config) on this rank's GPU
    full = make_head(d, synth.default_config(), params, a.precision, device=dev); set_bank(full, bank, dev)
    Lf, gtf, gvf, ggt, ggv = run(full, cat("text_feat"), cat("video_feat"), cat("text_mask"), cat("video_mask"),
                                 cat("idx"), cat("global_text"), cat("global_video"))
    # sharded head on the local slice
    cfg = synth.default_config(world_size=world, local_rank=local, rank=rank)
    sh = make_head(d, cfg, params, a.precision, device=dev); set_bank(sh, bank, dev)
    me = parts[rank].to(dev)
    Ls, gts, gvs, sgt, sgv = run(sh, me.text_feat, me.video_feat, me.text_mask, me.video_mask, me.idx,
                                 me.global_text, me.global_video)
    tol = 2e-4 if a.precision == "fp32" else 2e-2      # bf16: split-K / atomics order differs between the two runs
    sl = slice(rank * b, (rank + 1) * b)
    errs = {
        "loss": float(((Ls - Lf).abs() / Lf.abs()).max()),
        "text": rel_l2(gts, gtf[sl]), "video": rel_l2(gvs, gvf[sl]), "gt": rel_l2(sgt, ggt[sl]), "gv": rel_l2(sgv, ggv[sl]),
        "w1": rel_l2(sh.text_weight_fc[0].weight.grad, full.text_weight_fc[0].weight.grad),
        "vw2": rel_l2(sh.video_weight_fc[2].weight.grad, full.video_weight_fc[2].weight.grad),
        "ls": rel_l2(sh.clip.logit_scale.grad, full.clip.logit_scale.grad),
    }
    ok = errs["loss"] < (1e-5 if a.precision == "fp32" else 1e-3) and all(v < tol for k, v in errs.items() if k != "loss")
    ok = ok and torch.equal(sh.mb_ind, full.mb_ind) and torch.equal(sh.mb_feat_t, full.mb_feat_t)
    n1 = sh.last_neighbors[0].cpu(); n1f = full.last_neighbors[0][sl].cpu()
    if a.precision == "fp32":
        ok = ok and torch.equal(n1, n1f)
    # ---- column-sharded evaluation vs the single-process matrix + compute_metrics
    from neighborretr_b200.evaluator import sharded_retrieval, similarity_matrix
    from neighborretr_b200.metrics import RetrievalMetrics
    ev = synth.make_batch(101, nt, nv, d=d, seed=77).to(dev)            # 101: ragged last shard
    full.eval()
    S = similarity_matrix(full, ev.text_mask, ev.video_mask, ev.text_feat, ev.video_feat)
    m_t2v, m_v2t = RetrievalMetrics.compute_metrics(S), RetrievalMetrics.compute_metrics(S.t().contiguous())
    s_t2v, s_v2t, (tv, ti) = sharded_retrieval(full, ev.text_mask, ev.video_mask, ev.text_feat, ev.video_feat, topk=10)
    ref_i = torch.sort(S, dim=1, descending=True, stable=True)[1][:, :10]
    ok_eval = s_t2v == m_t2v and s_v2t == m_v2t and torch.equal(ti.long(), ref_i)
    if a.precision != "fp32":      # ranks from the contraction's epilogue: not even the shard's block of S exists
        f_t2v, f_v2t, none = sharded_retrieval(full, ev.text_mask, ev.video_mask, ev.text_feat, ev.video_feat, fused=True)
        ok_eval = ok_eval and f_t2v == m_t2v and f_v2t == m_v2t and none is None
    errs["eval"] = 0.0 if ok_eval else 1.0
    ok = ok and ok_eval
    print(f"rank {rank}/{world} {'OK' if ok else 'FAIL'} {errs}", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
