#!/usr/bin/env python
"""Benchmark of the retrieval head: head fwd+bwd steps/s (BASELINE.json metric) on synthetic
MSR-VTT-shaped embeddings, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape msrvtt|activitynet]

A step = gather (N>1) + local/global similarity + the four losses + backward to the features, the token-weight
MLPs and logit_scale + memory-bank FIFO update, on one batch of b=128 samples per GPU.  `value` times steps with
inputs resident in HBM; `e2e` times the same step through the public module API from pinned HOST buffers with
the H2D copies and the D2H read of the losses inside the timed region.  L2 is flushed between timed steps.
`--impl reference` times the oracle port of the reference head (oracle/head.py) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from neighborretr_b200 import synth  # noqa: E402

B_PER_GPU = 128
D = 512


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    except Exception:
        return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback(B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture
K1_TRAFFIC = {(1, "msrvtt", "nr_maxsim2_fwd"): 23.73e6 + 0.44e6}


def flops_maxsim(rx, ry, nt, nv, d=D):
    return 2.0 * rx * ry * nt * nv * d


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(shape, steps, warmup, b=None):
    b = b or B_PER_GPU
    from oracle import head as O
    nt, nv, mrows = synth.SHAPES[shape]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    h = synth.make_batch(b, nt, nv, d=D, seed=1234)
    bank = synth.make_bank(mrows, nt, nv, d=D)
    params = synth.make_mlp_params(d=D)
    cfg = synth.default_config()
    lsp = torch.tensor(float(torch.log(torch.tensor(100.0))), requires_grad=True)

    def step():
        text = h.text_feat.clone().requires_grad_(True)
        video = h.video_feat.clone().requires_grad_(True)
        gt = h.global_text.clone().requires_grad_(True)
        gv = h.global_video.clone().requires_grad_(True)
        p = {k: {n: v.clone().requires_grad_(True) for n, v in sd.items()} for k, sd in params.items()}
        losses = O.compute_losses(text, video, h.text_mask, h.video_mask, bank.mb_feat_t, bank.mb_feat_v,
                                  bank.mb_mask_t, bank.mb_mask_v, gt, gv, p, lsp.exp(), cfg)
        losses[0].backward()
        return float(losses[0])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps * 1e3, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the reference evaluates the WHOLE global batch on every rank, so the job at N GPUs is one head
    # over B = 128*N samples: that is what the host cores are timed on, once (rank 0), with all their threads
    world = max(1, args.gpus)
    b_global = B_PER_GPU * world
    steps = max(1, min(args.steps, 8 if world == 1 else (4 if world == 2 else 2)))
    warm = max(1, min(args.warmup, 2 if world == 1 else 1))
    sps, ms, threads = cpu_reference_steps(args.shape, steps, warm, b=b_global)
    line = {
        "impl": "reference", "metric": "head_fwd_bwd_steps_per_s", "value": sps, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.shape, world),
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} fwd+bwd steps (after {warm} warm-up) of oracle/head.py compute_losses "
                                   f"(torch CPU, fp32) on the global batch B={b_global} in one process (the gathered "
                                   f"batch every reference rank evaluates; no collective)"},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def eval_cpu_baseline(nq=1000):
    """Reference eval path on the host: 64x64-tiled local_level + compute_metrics x2 (oracle port)."""
    from oracle import head as O
    from oracle import metrics as OM
    nt, nv, _ = synth.SHAPES["msrvtt"]
    torch.set_num_threads(os.cpu_count() or 1)
    h = synth.make_batch(nq, nt, nv, d=D, seed=77)
    params = synth.make_mlp_params(d=D)
    t0 = time.perf_counter()
    sim, sim_t = O.eval_similarity(h.text_feat, h.video_feat, h.text_mask, h.video_mask, params)
    r1 = OM.compute_metrics(sim); r2 = OM.compute_metrics(sim_t)
    return (time.perf_counter() - t0) * 1e3, r1["R1"], r2["R1"]


def run_eval_bench(model, dev, steps, nq=1000):
    """Eval sim + R@K latency (BASELINE.json configs[3]): 1k x 1k MSR-VTT-shaped test set, t2v and v2t."""
    from neighborretr_b200 import ops
    from neighborretr_b200.evaluator import _run_on_single_gpu, retrieval_metrics_fused, similarity_matrix
    from neighborretr_b200.metrics import RetrievalMetrics, metrics_from_counts
    nt, nv, _ = synth.SHAPES["msrvtt"]
    host = synth.make_batch(nq, nt, nv, d=D, seed=77)
    pin = {f: getattr(host, f).pin_memory() for f in ("text_feat", "video_feat", "text_mask", "video_mask")}
    res = {k: v.to(dev) for k, v in pin.items()}
    model.eval()

    def resident():
        s = similarity_matrix(model, res["text_mask"], res["video_mask"], res["text_feat"], res["video_feat"])
        g1, e1 = ops.rank_counts(s)
        g2, e2 = ops.rank_counts(s.t().contiguous())
        c = torch.stack([g1, e1, g2, e2]).cpu().numpy()
        return metrics_from_counts(c[0], c[1]), metrics_from_counts(c[2], c[3])

    def e2e():      # the reference-facing calls: host features in, numpy sim out, metrics dicts
        d = {k: v.to(dev, non_blocking=True) for k, v in pin.items()}
        sim, sim_t = _run_on_single_gpu(model, d["text_mask"], d["video_mask"], d["text_feat"], d["video_feat"])
        return RetrievalMetrics.compute_metrics(sim), RetrievalMetrics.compute_metrics(sim_t)

    def fused():    # ranks straight from the contraction's epilogue: the similarity matrix is never written
        return retrieval_metrics_fused(model, res["text_mask"], res["video_mask"], res["text_feat"], res["video_feat"])

    # ranks of the fused path against ranks counted on the materialised matrix, same token weights (two evaluations of
    # the weight MLPs differ in the last bit: float atomics in the second layer)
    from neighborretr_b200.evaluator import _eval_operands
    with torch.no_grad():
        tm_, vm_, tw_, vw_, prec_ = _eval_operands(model, res["text_mask"], res["video_mask"], res["text_feat"],
                                                   res["video_feat"], 64)
        s_, _ = ops.maxsim(res["text_feat"], res["video_feat"], tw_, vw_, tm_, vm_, prec_)
        fr = ops.FusedRanker(res["text_feat"], res["video_feat"], tw_, vw_, tm_, vm_, prec_)
        cf = torch.stack(fr.counts(fr.diagonal(nq)))
        cm = torch.stack([*ops.rank_counts(s_), *ops.rank_counts(s_.t().contiguous())])
    out = {"fused_equals_materialised": bool(torch.equal(cf, cm))}       # every count of every query, both directions
    for name, fn in (("fused_resident_ms", fused), ("resident_ms", resident), ("e2e_ms", e2e)):
        for _ in range(3):
            r = fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = fn()
        torch.cuda.synchronize()
        out[name] = (time.perf_counter() - t0) * 1e3 / steps
    out.update({"queries": nq, "gallery": nq, "t2v_R1": r[0]["R1"], "v2t_R1": r[1]["R1"], "unit": "ms",
                "what": "similarity matrix (token-weight MLPs + max-sim) + t2v and v2t R@1/5/10/50, MdR, MeanR",
                "flops": flops_maxsim(nq, nq, nt, nv)})
    model.train()
    return out


def run_eval_workload(args):
    """BASELINE.json configs[4], evaluation half: N x N test set (default 100 000) with the video gallery sharded by
    column over the ranks (evaluator.sharded_retrieval): per-rank similarity block S[:, cols_r] on the tensor cores,
    per-shard rank counts all-reduced, per-shard top-10 merged.  One "step" = one full evaluation (t2v and v2t
    metrics + merged top-10), features resident in HBM; `e2e` repeats it with the features copied from pinned host
    memory inside the timed region.  --impl reference: the oracle's tiled evaluation on the host cores over a bounded
    sample, scaled by the number of query x gallery pairs."""
    import torch.distributed as dist
    n = args.eval_size
    nt, nv, _ = synth.SHAPES[args.shape]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    cfg_line = {"workload": f"{args.shape}_eval_{n}x{n}_column_sharded", "queries": n, "gallery": n, "words": nt,
                "frames": nv, "dim": D, "topk": 10, "parallelism": f"column shards x{world}",
                "l2": "inputs (GBs of features) larger than L2"}
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import head as O
        from oracle import metrics as OM
        ns = 1000
        torch.set_num_threads(os.cpu_count() or 1)
        h = synth.make_batch(ns, nt, nv, d=D, seed=77)
        params = synth.make_mlp_params(d=D)
        t0 = time.perf_counter()
        sim, sim_t = O.eval_similarity(h.text_feat, h.video_feat, h.text_mask, h.video_mask, params)
        OM.compute_metrics(sim); OM.compute_metrics(sim_t)
        ms = (time.perf_counter() - t0) * 1e3 * (n / ns) ** 2
        print(json.dumps({"impl": "reference", "metric": "eval_sim_rank_latency_ms", "value": ms, "unit": "ms",
                          "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": ms, "higher_is_better": False,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_line,
                          "cpu_baseline": {"value": ms, "unit": "ms", "cores": os.cpu_count(), "kind": "port",
                                           "sample": f"{ns} x {ns} tiled evaluation + both metric passes, scaled by "
                                                     f"({n}/{ns})^2 pairs"},
                          "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from neighborretr_b200 import ops, selfcheck
    from neighborretr_b200.evaluator import sharded_retrieval
    model = selfcheck.make_model(synth.default_config(world_size=world, local_rank=local, rank=rank), dev, args.precision).eval()
    # every rank holds all features, as after the reference's eval gather (training/evaluator.py:173-189)
    g = torch.Generator().manual_seed(77)
    chunk = 10000
    tf = torch.empty(n, nt, D, device=dev); vf = torch.empty(n, nv, D, device=dev)
    tm = torch.empty(n, nt, dtype=torch.int64, device=dev); vm = torch.empty(n, nv, dtype=torch.int64, device=dev)
    for c0 in range(0, n, chunk):
        hb = synth.make_batch(min(chunk, n - c0), nt, nv, d=D, seed=77 + c0)
        tf[c0:c0 + chunk] = hb.text_feat.to(dev); vf[c0:c0 + chunk] = hb.video_feat.to(dev)
        tm[c0:c0 + chunk] = hb.text_mask.to(dev); vm[c0:c0 + chunk] = hb.video_mask.to(dev)

    fused = args.eval_ranks == "fused"
    # how the ranks are produced is a property of OUR arm, not of the workload both arms share: own top-level key
    ranks_how = ("counted in the contraction's epilogue, S never written (no top-k lists)" if fused else
                 "per-rank block of S written, rank-count + top-10 kernels over it")

    def once():
        return sharded_retrieval(model, tm, vm, tf, vf, topk=10, fused=fused)

    def sync_all():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 2))):
        r = once()
    steps = max(1, min(args.steps, 5))
    ops.LAUNCHES["count"] = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    evs = []
    for _ in range(steps):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record(); r = once(); e_.record()
        evs.append((s_, e_))
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    launches = ops.LAUNCHES["count"]
    t = torch.tensor([sum(a.elapsed_time(b_) for a, b_ in evs)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    if rank == 0:
        pk = peaks()
        fl = flops_maxsim(n, n, nt, nv) / world                      # per rank: its column shard
        ach = fl / (ms * 1e-3) / 1e12
        print(json.dumps({
            "metric": "eval_sim_rank_latency_ms", "value": ms, "unit": "ms", "n_gpus": world, "steps": steps,
            "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else args.precision, "data": "synthetic",
            "config": cfg_line, "eval_ranks": ranks_how, "t2v_R1": r[0]["R1"], "v2t_R1": r[1]["R1"],
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"kernel": "nr_maxsim2_fwd", "bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"],
                         "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"], "traffic": None,
                         "note": "per-rank similarity flops 2*Q*(N/W)*Nt*Nv*D / the WHOLE evaluation call (MLPs, "
                                 "preparation, contraction, rank counts, top-k, collectives)"}}), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


def workload_config(shape, world):
    """The workload both arms run (the arithmetic type of an arm is its `dtype`, not part of the workload)."""
    nt, nv, mrows = synth.SHAPES[shape]
    return {"workload": f"{shape}_head_b{B_PER_GPU}_per_gpu", "per_gpu_batch": B_PER_GPU,
            "global_batch": B_PER_GPU * world, "words": nt, "frames": nv, "dim": D, "memory_rows": mrows,
            "num_neighbors": 20, "sinkhorn_iters": 50,
            "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"dp{world}"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from neighborretr_b200 import ops
    from neighborretr_b200.modeling import NeighborRetr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nt, nv, mrows = synth.SHAPES[args.shape]
    cfg = synth.default_config(world_size=world, local_rank=local, rank=rank)
    model = NeighborRetr(cfg, width=D)
    for name, sd in synth.make_mlp_params(d=D).items():
        getattr(model, name).load_state_dict(sd)
    model.clip.logit_scale.data.fill_(float(torch.log(torch.tensor(100.0))))
    model.head_precision = args.precision
    model.head_bwd_precision = args.bwd_precision or args.precision
    model = model.to(dev).train()
    bank = synth.make_bank(mrows, nt, nv, d=D)
    host = synth.make_batch(B_PER_GPU, nt, nv, d=D, seed=1234, rank=rank)
    pinned = {f: getattr(host, f).pin_memory() for f in host.__dataclass_fields__}
    resident = {k: v.to(dev) for k, v in pinned.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss_host = torch.empty(5, dtype=torch.float32).pin_memory()

    def reset_bank():
        model.mb_ind = bank.mb_ind.to(dev); model.mb_feat_t = bank.mb_feat_t.to(dev)
        model.mb_feat_v = bank.mb_feat_v.to(dev); model.mb_mask_t = bank.mb_mask_t.to(dev)
        model.mb_mask_v = bank.mb_mask_v.to(dev); model.mb_batch = mrows

    reset_bank()

    def step(src, from_host):
        if from_host:
            t = {k: v.to(dev, non_blocking=True) for k, v in src.items()}
        else:
            t = src
        text = t["text_feat"].detach().requires_grad_(True)
        video = t["video_feat"].detach().requires_grad_(True)
        gt = t["global_text"].detach().requires_grad_(True)
        gv = t["global_video"].detach().requires_grad_(True)
        model.zero_grad(set_to_none=True)
        losses = model.head_forward(text, video, t["text_mask"], t["video_mask"], t["idx"], global_feats=(gt, gv))
        losses[0].backward()
        if from_host:
            loss_host.copy_(torch.stack([x.detach() for x in losses]), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return losses

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        evs = []
        sync_all()
        for _ in range(steps):
            flush_buf.fill_(1)                         # L2 flush, outside the timed window
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        sync_all()
        tot = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([tot], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(resident, False)
    use_graph = not args.no_graph
    gstep = None
    parity = None
    do_check = (world > 1) if args.check is None else bool(args.check)
    if use_graph:
        from neighborretr_b200.graph import FIELDS, GraphedHeadStep
        reset_bank()
        res_list = [resident[f] for f in FIELDS]
        pin_list = [pinned[f] for f in FIELDS]
        gstep = GraphedHeadStep(model, res_list, warmup=warm)
        if do_check:
            # the graph-captured step that is timed below, against the single-process full-batch head
            from neighborretr_b200 import selfcheck
            ref = selfcheck.full_batch_reference(dev, world, args.shape, B_PER_GPU, mrows, args.precision,
                                                 args.bwd_precision)
            out = gstep(*res_list)
            err = selfcheck.compare_step(ref, rank, B_PER_GPU, out, gstep.grads, model)
            ok, parity = selfcheck.check_all_ranks(err, args.precision, dev)
            parity["step"] = "GraphedHeadStep replay (captured NCCL collectives)" if world > 1 else "GraphedHeadStep replay"
            del ref
            if not ok:
                raise SystemExit(f"bench.py --check FAILED on rank {rank}: {err}")
            gstep.set_bank(bank)
        run_value = lambda: gstep(*res_list)
        run_e2e_serial = lambda: gstep(*pin_list, sync_losses_to=loss_host)

        def run_e2e():
            # steady-state input pipeline: this step consumes the batch staged by the previous call, then the H2D copy
            # of the NEXT step's pinned batch is started and runs under this step's replay (one H2D of the whole
            # batch and one D2H read of the losses per step, both inside the timed region)
            return gstep(sync_losses_to=loss_host, prefetched=True, prefetch_next=pin_list)

        gstep.prefetch(*pin_list)               # fill the pipeline: the first timed step finds its batch staged
    else:
        if do_check:
            from neighborretr_b200 import selfcheck
            ref = selfcheck.full_batch_reference(dev, world, args.shape, B_PER_GPU, mrows, args.precision,
                                                 args.bwd_precision)
            reset_bank()
            t = resident
            leaves = {f: t[f].detach().clone().requires_grad_(True)
                      for f in ("text_feat", "video_feat", "global_text", "global_video")}
            model.zero_grad(set_to_none=True)
            ls_ = model.head_forward(leaves["text_feat"], leaves["video_feat"], t["text_mask"], t["video_mask"], t["idx"],
                                     global_feats=(leaves["global_text"], leaves["global_video"]))
            ls_[0].backward()
            err = selfcheck.compare_step(ref, rank, B_PER_GPU, torch.stack([x.detach() for x in ls_]),
                                         {f: v.grad for f, v in leaves.items()}, model)
            ok, parity = selfcheck.check_all_ranks(err, args.precision, dev)
            parity["step"] = "eager head_forward + backward"
            del ref
            if not ok:
                raise SystemExit(f"bench.py --check FAILED on rank {rank}: {err}")
            reset_bank()
        run_value = lambda: step(resident, False)
        run_e2e = run_e2e_serial = lambda: step(pinned, True)
    for _ in range(warm):
        run_value()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.LAUNCHES["count"] = 0
    ms_total = timed(run_value, args.steps)
    launches = ops.LAUNCHES["count"]
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        run_e2e()
    ms_e2e = timed(run_e2e, args.steps)
    for _ in range(2):
        run_e2e_serial()
    ms_e2e_serial = timed(run_e2e_serial, args.steps)
    # eager module-API step from host buffers (no graph), and per-launch CUDA-event timing of the dominant kernel
    for _ in range(2):
        step(pinned, True)
    ms_e2e_eager = timed(lambda: step(pinned, True), args.steps)
    kname = "nr_maxsim2_fwd" if (args.precision == "bf16" and (args.bwd_precision or "bf16") == "bf16") else "nr_maxsim_fwd"
    ops.KERNEL_TIMER.enable(kname)
    ksteps = min(args.steps, 10)

    def held_step():
        # park the stream on a spin kernel so that the host is ahead of the device when the timed launch is enqueued:
        # the event pair then brackets GPU time only (see ops._KernelTimer)
        ops.KERNEL_TIMER.hold(4.0)
        step(resident, False)

    timed(held_step, ksteps)
    kt = ops.KERNEL_TIMER.collect()
    ops.KERNEL_TIMER.disable()
    # ---- the module API a trainer uses (model.head_forward(...) -> loss.backward()) with head_graph on: one graph
    #      replay per step behind autograd (graph.GraphedHead), host buffers in, losses read back
    extra = {}
    if not args.no_extra:
        reset_bank()
        model.head_graph = True
        for _ in range(3):
            step(pinned, True)
        ms_api = timed(lambda: step(pinned, True), args.steps)
        # ... and as the reference's trainer sees it: the features come out of the encoders ON the device, the losses
        # stay device tensors until the trainer logs them (training/trainer.py:84-125) — the cost of the API itself
        for _ in range(2):
            step(resident, False)
        ms_api_res = timed(lambda: step(resident, False), args.steps)
        model.head_graph = False
        extra["module_api_graph_steps_per_s"] = args.steps / (ms_api * 1e-3)
        extra["module_api_graph_resident_steps_per_s"] = args.steps / (ms_api_res * 1e-3)
        extra["module_api_note"] = ("model.head_forward(...) + loss.backward() with head_graph=True: from pinned host "
                                    "buffers with the losses read back (serial H2D of 10 MB in every step), and with "
                                    "device-resident features as the reference trainer has them")
        # ---- the other arithmetic modes of the same step (graph replay, inputs resident)
        from neighborretr_b200 import selfcheck
        from neighborretr_b200.graph import FIELDS as _F, GraphedHeadStep as _G
        for mode in ("bf16x3", "fp32"):
            if mode == args.precision or (mode == "fp32" and (world > 1 or args.shape != "msrvtt" or B_PER_GPU > 256)):
                continue
            m2 = selfcheck.make_model(cfg, dev, mode)
            selfcheck.set_bank(m2, bank, dev)
            g2 = _G(m2, [resident[f] for f in _F], warmup=3)
            rl = [resident[f] for f in _F]
            for _ in range(2):
                g2(*rl)
            nst = max(3, args.steps // 2)
            ms2 = timed(lambda: g2(*rl), nst)
            extra[mode] = {"steps_per_s": nst / (ms2 * 1e-3), "ms_per_step": ms2 / nst,
                           "losses": [float(x) for x in g2.losses.tolist()]}
            if world > 1:
                g2.graph.reset()
            del g2, m2

    def finish():
        """Multi-rank exit: captured NCCL work keeps the communicator busy, and destroy_process_group() then
        blocks; release the graph, synchronise, and leave without tearing NCCL down."""
        if world > 1:
            nonlocal gstep
            if gstep is not None:
                gstep.graph.reset()
                gstep = None
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return

    B = B_PER_GPU * world
    pk = peaks()
    # dominant kernel = nr_maxsim2_fwd: ONE launch per step holding the batch pair(s) and both bank pairs, both
    # max directions from the same accumulator tile.  Algorithmic flops of a rank-step's forward contractions
    # (SURVEY.md §8(d)): every S entry and every bank entry counted ONCE globally, un-padded tokens only:
    #   2*Nt*Nv*D * (b*B + 2*b*M)   with b = per-rank rows.
    # Executed MMA work: x 128/120 * 256/240 tile padding at 24/12 tokens; for W > 1 a rank contracts only its text rows
    # (the column block arrives by an all-to-all of [b,b] similarity blocks), so executed = algorithmic there too.
    flops_step = flops_maxsim(B_PER_GPU, B, nt, nv) + 2 * flops_maxsim(B_PER_GPU, mrows, nt, nv)
    n_l = max(kt["launches"], 1)
    achieved = flops_step * ksteps / (kt["ms"] * 1e-3) / 1e12 if kt["ms"] > 0 else 0.0
    peak = pk["bf16_sustained"]
    line = {
        "metric": "head_fwd_bwd_steps_per_s", "value": args.steps / (ms_total * 1e-3), "unit": "steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(args.shape, world), "head_precision": args.precision,
        "samples_per_s": args.steps * B / (ms_total * 1e-3),
        "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 20, "ms_per_step": ms_e2e / args.steps,
                "api": ("GraphedHeadStep: prefetch(next pinned host batch) + replay; the H2D copy of step i+1 overlaps "
                        "step i, the losses are read back and synchronised every step") if use_graph
                       else "model.head_forward + backward",
                "serial_steps_per_s": args.steps / (ms_e2e_serial * 1e-3),
                "serial_note": "same call without prefetch: H2D copy, replay, D2H read strictly one after the other",
                "eager_module_api_steps_per_s": args.steps / (ms_e2e_eager * 1e-3)},
        "cuda_graph": bool(use_graph),
        "parity_checked": parity,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": K1_TRAFFIC.get((world, args.shape, kname)),
                     "traffic_note": "a constant, not measured by this run: dram bytes read+written per launch in the ncu "
                                     "--set full capture of tools/k2_only.py (profiles/r2_k2_fwd_ncu_full.txt): the 23.6e6 B "
                                     "of bf16 operands are read from HBM once, the 27e6 B of saved max/arg-max and "
                                     "similarities stay in L2",
                     "launches_timed": n_l, "avg_launch_ms": kt["ms"] / n_l,
                     "share_of_step": (kt["ms"] / ksteps) / (ms_total / args.steps) if ms_total else None,
                     "timed_in": "eager steps on the launching stream (events cannot be read inside a graph replay), the "
                                 "stream parked on a spin kernel first so that the event pair brackets device time only",
                     "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside the step)"},
    }
    if world == 1:
        line["eval"] = run_eval_bench(model, dev, 10)
        ev = line["eval"]
        ach = ev["flops"] / (ev["resident_ms"] * 1e-3) / 1e12
        ev["roofline"] = {"kernel": kname, "bound": "tensor", "achieved": ach, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                          "frac": ach / pk["bf16_burst"],
                          "note": "2*Q*N*Nt*Nv*D flops of the similarity matrix / the whole resident eval call (token-weight "
                                  "MLPs, preparation, max-sim, two rank-count passes, D2H of the counts); peak = burst "
                                  "bf16 (a ~1 ms call)"}
    line["modes"] = extra
    if world == 1 and not args.no_cpu_baseline:
        ems, er1, er2 = eval_cpu_baseline()
        line["eval"]["cpu_baseline_ms"] = ems
        line["eval"]["cpu_t2v_R1"] = er1
        sps, ms, threads = cpu_reference_steps(args.shape, 4, 1)
        line["cpu_baseline"] = {"value": sps, "unit": "steps/s", "cores": threads, "kind": "port",
                                "sample": "4 fwd+bwd steps (after 1 warm-up) of oracle/head.py compute_losses on the "
                                          f"same {args.shape} b={B_PER_GPU} batch, torch CPU fp32"}
    print(json.dumps(line), flush=True)
    finish()


def main():
    global B_PER_GPU
    if os.environ.get("NR_DEBUG_HANG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["NR_DEBUG_HANG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="msrvtt", choices=list(synth.SHAPES))
    ap.add_argument("--precision", default=os.environ.get("NR_HEAD_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--bwd-precision", default=os.environ.get("NR_HEAD_BWD_PRECISION"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eager module API instead of the CUDA graph")
    ap.add_argument("--check", dest="check", action="store_const", const=1, default=None,
                    help="before timing, compare the (graph-captured) step with the single-process full-batch head "
                         "(default: on when N > 1)")
    ap.add_argument("--no-check", dest="check", action="store_const", const=0)
    ap.add_argument("--per-gpu-batch", type=int, default=128,
                    help="samples per GPU (BASELINE configs[4]: 1024 per GPU on 8 GPUs = global batch 8192)")
    ap.add_argument("--workload", default="head", choices=["head", "eval"],
                    help="head: fwd+bwd steps/s (the metric BASELINE.json quotes); eval: column-sharded similarity + "
                         "R@K latency on an --eval-size x --eval-size test set (configs[4]: 100000)")
    ap.add_argument("--eval-size", type=int, default=100000)
    ap.add_argument("--eval-ranks", default="fused", choices=["fused", "materialised"],
                    help="--workload eval: rank counts from the contraction's epilogue (no similarity matrix), or the "
                         "per-rank block of S + rank-count / top-k kernels")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra precision-mode / module-API measurements")
    args = ap.parse_args()
    B_PER_GPU = args.per_gpu_batch
    if args.workload == "eval":
        run_eval_workload(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
        run_ours(args)


if __name__ == "__main__":
    main()
