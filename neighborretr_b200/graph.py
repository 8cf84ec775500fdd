"""Whole-step CUDA graph of the retrieval head: head_forward (losses) + backward (gradients of the features, the
global features, the token-weight MLPs and logit_scale) + memory-bank FIFO update, captured once for fixed shapes
and replayed with one launch.  At b=128 the head is launch-bound (~100 small kernels per step), so the graph is
what turns GPU time into steps/s.

    step = GraphedHeadStep(model, example_batch)            # model: neighborretr_b200.modeling.NeighborRetr (cuda)
    losses = step(text_feat, video_feat, text_mask, video_mask, idx, global_text, global_video)
    step.grads["text_feat"], model.text_weight_fc[0].weight.grad, ...   # static tensors, overwritten per replay

Inputs may live in (pinned) host memory: they are copied into the static device buffers on the replay stream.
The memory bank is advanced IN PLACE inside the graph (cat(new, old)[:capacity], reference modeling.py:235-249),
so `model.mb_*` keep their storage across replays.  Single-GPU only (world_size == 1).
"""
from __future__ import annotations

import torch

from . import ops

FIELDS = ("text_feat", "video_feat", "text_mask", "video_mask", "idx", "global_text", "global_video")
GRAD_FIELDS = ("text_feat", "video_feat", "global_text", "global_video")


class GraphedHeadStep:
    def __init__(self, model, example, warmup=3):
        self.world = getattr(model.config, "world_size", 1)
        self.model = model
        dev = next(model.parameters()).device
        self.static = {}
        for f, t in zip(FIELDS, example):
            s = t.detach().to(dev).clone()
            if f in GRAD_FIELDS:
                s.requires_grad_(True)
            self.static[f] = s
        self.bank_names = ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v")
        for n in self.bank_names:                       # own the bank storage
            setattr(model, n, getattr(model, n).detach().to(dev).clone())
        bank0 = {n: getattr(model, n).clone() for n in self.bank_names}
        self.params = [p for p in model.parameters() if p.requires_grad]
        self._e0 = torch.tensor([1.0, 0.0, 0.0, 0.0, 0.0], device=dev)
        self._fifo_stream = torch.cuda.Stream(device=dev)
        self._stage = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._zero_grads()
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._zero_grads()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCHES["count"]
        # thread_local: other threads (the NCCL watchdog polling events) may keep making CUDA calls during capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.losses = self._body()
        self.launches_per_replay = ops.LAUNCHES["count"] - n0
        for n in self.bank_names:                       # warm-up steps advanced the bank: restore it in place
            getattr(model, n).copy_(bank0[n])
        self.grads = {f: self.static[f].grad for f in GRAD_FIELDS}

    def _zero_grads(self):
        for t in list(self.static.values()) + self.params:
            t.grad = None

    def _body(self):
        m, s = self.model, self.static
        cfg = m.config
        if self.world == 1:
            logit_scale = m.clip.logit_scale.exp()
            losses = m._compute_losses(s["text_feat"], s["video_feat"], s["text_mask"], s["video_mask"], m.mb_feat_t,
                                       m.mb_feat_v, m.mb_mask_t, m.mb_mask_v, cfg.centrality_scale, cfg.beta,
                                       cfg.num_neighbors, cfg.temperature, logit_scale,
                                       global_feats=(s["global_text"], s["global_video"]))
            new_rows = (s["idx"], s["text_feat"], s["video_feat"], s["text_mask"], s["video_mask"])
        else:       # row-block sharded head; the NCCL collectives are captured in the graph
            from .until_module import _gather_contiguous
            losses, (ta, va, tma, vma) = m._sharded_losses(s["text_feat"], s["video_feat"], s["text_mask"],
                                                           s["video_mask"], (s["global_text"], s["global_video"]))
            new_rows = (_gather_contiguous(s["idx"], self.world), ta, va, tma, vma)
        # With bf16 weight-MLP GEMMs the backward reads bf16 copies, never the bank itself: the FIFO update can then
        # leave the critical path and run on its own branch next to the backward.
        early_fifo = self.world == 1 and m._mlp_precision() == "bf16"
        if early_fifo:
            main = torch.cuda.current_stream()
            self._fifo_stream.wait_stream(main)
            with torch.cuda.stream(self._fifo_stream):
                self._fifo(new_rows)
        out5 = getattr(m, "last_out5", None) if self.world == 1 else None
        if out5 is not None and out5.requires_grad:
            # d total / d out5 = e0: skips the unbind/stack bookkeeping kernels of losses[0].backward()
            out5.backward(gradient=self._e0)
            m.last_out5 = None
        else:
            out5 = None
            losses[0].backward()
        if early_fifo:
            main.wait_stream(self._fifo_stream)
        else:
            if self.world > 1:
                m.wait_gathered_text()
            self._fifo(new_rows)
        return out5.detach() if out5 is not None else torch.stack([x.detach() for x in losses])

    def _fifo(self, new_rows):
        """bank <- cat(new, bank)[:capacity] in place (reference modeling.py:235-249)."""
        m = self.model
        with torch.no_grad():
            cap = m.mb_feat_v.shape[0]
            for name, new in zip(("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v"), new_rows):
                bank = getattr(m, name)
                bank.copy_(ops.fifo_update(new.detach().to(bank.dtype), bank, cap))

    # ---- input prefetch: the H2D copy of step i+1 overlaps the replay of step i --------------------------------
    def prefetch(self, *batch):
        """Start copying a batch (pinned host or device tensors) into the staging buffer on a copy stream; a later
        __call__(prefetched=True) moves it into the static inputs with a device-to-device copy.  The copy waits until
        the previous staging content has been consumed; use __call__(..., prefetch_next=batch) inside a loop."""
        dev = self.static[FIELDS[0]].device
        if self._stage is None:
            self._stage = {f: torch.empty_like(self.static[f].data) for f in FIELDS}
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free)       # the previous staging content has been consumed
            for f, t in zip(FIELDS, batch):
                self._stage[f].copy_(t, non_blocking=True)
            self._staged.record(self._copy_stream)

    def __call__(self, *batch, sync_losses_to=None, prefetched=False, prefetch_next=None):
        """Copy the batch into the static buffers (H2D if it lives on the host), replay, return the static
        [total, centrality, uniform, neighbor, kl] tensor (or copy it into the pinned host tensor given).
        prefetched=True: take the batch a previous prefetch() staged on the device instead; prefetch_next=batch
        starts staging the following step's batch as soon as the staging buffer has been consumed, i.e. its H2D
        copy runs under this step's replay."""
        if prefetched:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._staged)
            for f in FIELDS:
                self.static[f].data.copy_(self._stage[f], non_blocking=True)
            self._stage_free.record(cur)
            if prefetch_next is not None:
                self.prefetch(*prefetch_next)
        else:
            for f, t in zip(FIELDS, batch):
                self.static[f].data.copy_(t, non_blocking=True)
        self.graph.replay()
        ops.LAUNCHES["count"] += self.launches_per_replay
        if sync_losses_to is not None:
            sync_losses_to.copy_(self.losses, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return sync_losses_to
        return self.losses
