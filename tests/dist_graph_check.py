"""Multi-rank check of the CUDA-GRAPH captured sharded head step (GraphedHeadStep with world_size > 1: captured NCCL
gathers / all-to-all / reductions, deferred text gather on a side stream, in-place bank FIFO) against the
single-process full-batch head — the computation bench.py times at N > 1.  NCCL only (gloo collectives cannot be
captured), one GPU per rank:

    torchrun --nproc-per-node 2 tests/dist_graph_check.py [--precision bf16|fp32] [--b 32] [--replays 3]

Also replays the graph several times and checks the LAST replay against an eager sharded step from the same bank
state, so that stale static buffers / bank updates across replays would be caught."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--b", type=int, default=32)
    ap.add_argument("--shape", default="msrvtt")
    ap.add_argument("--mrows", type=int, default=96)
    ap.add_argument("--replays", type=int, default=3)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from neighborretr_b200 import selfcheck, synth
    from neighborretr_b200.graph import FIELDS, GraphedHeadStep
    nt, nv, _ = synth.SHAPES[a.shape]
    bank = synth.make_bank(a.mrows, nt, nv)
    ref = selfcheck.full_batch_reference(dev, world, a.shape, a.b, a.mrows, a.precision)
    cfg = synth.default_config(world_size=world, local_rank=local, rank=rank)
    model = selfcheck.make_model(cfg, dev, a.precision)
    selfcheck.set_bank(model, bank, dev)
    me = synth.make_batch(a.b, nt, nv, seed=1234, rank=rank).to(dev)
    batch = [getattr(me, f) for f in FIELDS]
    gstep = GraphedHeadStep(model, batch)
    losses = gstep(*batch)
    err = selfcheck.compare_step(ref, rank, a.b, losses, gstep.grads, model)
    ok, summary = selfcheck.check_all_ranks(err, a.precision, dev)
    # ---- further replays on fresh batches: the last one against an eager sharded step from the same bank state
    for i in range(1, a.replays):
        nxt = synth.make_batch(a.b, nt, nv, seed=777 + 10 * i, rank=rank).to(dev)
        nb = [getattr(nxt, f) for f in FIELDS]
        if i == a.replays - 1:
            eager = selfcheck.make_model(cfg, dev, a.precision)
            for n in gstep.bank_names:
                setattr(eager, n, getattr(model, n).clone())
            leaves = {f: getattr(nxt, f).clone().requires_grad_(True)
                      for f in ("text_feat", "video_feat", "global_text", "global_video")}
            le = eager.head_forward(leaves["text_feat"], leaves["video_feat"], nxt.text_mask, nxt.video_mask, nxt.idx,
                                    global_feats=(leaves["global_text"], leaves["global_video"]))
            le[0].backward()
            le = torch.stack([x.detach() for x in le])
        lg = gstep(*nb).clone()
    ok2 = True
    if a.replays > 1:
        ltol, gtol = selfcheck.tolerances(a.precision)
        e2 = {"loss": float(((lg - le).abs() / le.abs().clamp_min(1e-12)).max()),
              "text": selfcheck._rel_l2(gstep.grads["text_feat"], leaves["text_feat"].grad),
              "video": selfcheck._rel_l2(gstep.grads["video_feat"], leaves["video_feat"].grad),
              "w1": selfcheck._rel_l2(model.text_weight_fc[0].weight.grad, eager.text_weight_fc[0].weight.grad),
              "bank": float(not torch.equal(model.mb_feat_t, eager.mb_feat_t))}
        ok2 = e2["loss"] < ltol and max(e2["text"], e2["video"], e2["w1"]) < gtol and e2["bank"] == 0.0
        err["replay_vs_eager"] = e2
    print(f"rank {rank}/{world} {'OK' if ok and ok2 else 'FAIL'} {summary} {err}", flush=True)
    flag = torch.tensor([1.0 if ok and ok2 else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    good = flag.item() == 1.0
    gstep.graph.reset()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if good else 1)      # captured NCCL work keeps the communicator busy: leave without tearing it down


if __name__ == "__main__":
    main()
