"""Whole-step CUDA graph of the retrieval head: head_forward (losses) + backward (gradients of the features, the
global features, the token-weight MLPs and logit_scale) + memory-bank FIFO update, captured once for fixed shapes
and replayed with one launch.  At b=128 the head is launch-bound (~100 small kernels per step), so the graph is
what turns GPU time into steps/s.

    step = GraphedHeadStep(model, example_batch)            # model: neighborretr_b200.modeling.NeighborRetr (cuda)
    losses = step(text_feat, video_feat, text_mask, video_mask, idx, global_text, global_video)
    step.grads["text_feat"], model.text_weight_fc[0].weight.grad, ...   # static tensors, overwritten per replay

Inputs may live in (pinned) host memory: they are copied into the static device buffers on the replay stream.
The memory bank is advanced IN PLACE inside the graph (cat(new, old)[:capacity], reference modeling.py:235-249),
so `model.mb_*` keep their storage across replays.

world_size > 1: the body is the row-block sharded head (sharded.py); its NCCL collectives (gathers, the [b,b]
all-to-all, loss / gradient reductions) are captured in the same graph, so every rank must construct and replay
its GraphedHeadStep in lockstep.  Parity of the replayed step against the single-process full-batch head:
selfcheck.py (bench.py --check, tests/dist_graph_check.py).
"""
from __future__ import annotations

import os

import torch

from . import ops

FIELDS = ("text_feat", "video_feat", "text_mask", "video_mask", "idx", "global_text", "global_video")
GRAD_FIELDS = ("text_feat", "video_feat", "global_text", "global_video")


def head_params(model):
    """The parameters the head itself differentiates: the two token-weight MLPs and logit_scale."""
    return list(model.text_weight_fc.parameters()) + list(model.video_weight_fc.parameters()) + [model.clip.logit_scale]


def _slab_like(tensors, dev):
    """One uint8 device buffer holding a tensor of every given shape / dtype at 256-byte aligned offsets; returns
    (slab, {field: view})."""
    offs, total = [], 0
    for t in tensors:
        offs.append(total)
        total += (t.numel() * t.element_size() + 255) // 256 * 256
    slab = torch.empty(total, dtype=torch.uint8, device=dev)
    views = {}
    for f, t, o in zip(FIELDS, tensors, offs):
        views[f] = slab[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
    return slab, views


class GraphedHeadStep:
    def __init__(self, model, example, warmup=3, explicit_grads=False):
        """explicit_grads: capture torch.autograd.grad(...) instead of .backward(): the gradients of the four
        differentiable inputs and of head_params(model) are kept as static tensors (`grad_list`) and no .grad
        attribute is touched — what the trainer-compatible wrapper (GraphedHead) hands back to autograd."""
        self.world = getattr(model.config, "world_size", 1)
        if self.world > 1 and not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = 1
        self.model = model
        self.explicit = bool(explicit_grads)
        dev = example[0].device if example[0].is_cuda else next(model.parameters()).device
        # the static inputs are views of ONE byte slab: a staged batch moves in with a single device-to-device copy
        self._static_slab, self.static = _slab_like(example, dev)
        for f, t in zip(FIELDS, example):
            self.static[f].copy_(t.detach())
            if f in GRAD_FIELDS:
                self.static[f].requires_grad_(True)
        self.bank_names = ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v")
        bank0 = {n: getattr(model, n).detach().to(dev).clone() for n in self.bank_names}     # reference order
        self.ring = self._make_ring(model, bank0, self.static["text_feat"].shape[0])
        if self.ring is None:
            for n in self.bank_names:                   # own the bank storage
                setattr(model, n, bank0[n].clone())
        self.params = head_params(model)
        self.grad_list = None
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
        if self.ring is not None:                       # warm-up steps advanced the bank: restore it in place
            self.ring.load(*[bank0[n] for n in self.bank_names])
        else:
            for n in self.bank_names:
                getattr(model, n).copy_(bank0[n])
        self.grads = dict(zip(GRAD_FIELDS, self.grad_list[:len(GRAD_FIELDS)]))
        if not self.explicit:              # the documented surface: parameter .grad = the static gradient tensors
            for p_, g_ in zip(self.params, self.grad_list[len(GRAD_FIELDS):]):
                p_.grad = g_

    def _make_ring(self, model, bank0, b):
        """The persistent prepared bank (bank.BankRing) when the step runs the fused tensor-core kernels; None
        otherwise (fp32 mode, unsupported token counts, or NR_BANK_RING=0): the bank then stays five tensors that
        the captured FIFO kernels rewrite."""
        if os.environ.get("NR_BANK_RING", "1") == "0" or bank0["mb_feat_v"].dim() != 3 or bank0["mb_feat_v"].shape[0] == 0:
            return None
        prec, bprec = model._head_precision(), model._head_bwd_precision()
        nt, nv, d = bank0["mb_feat_t"].shape[1], bank0["mb_feat_v"].shape[1], bank0["mb_feat_t"].shape[2]
        x3 = prec == "bf16x3"
        if prec not in ("bf16", "bf16x3") or bprec != prec or not ops.USE_FUSED_MAXSIM \
                or not ops.maxsim2_supported(nt, nv, d * (3 if x3 else 1)):
            return None
        from .bank import BankRing
        mlp_bf16 = model._mlp_precision() == "bf16" and ops.USE_OWN_GEMM
        rows = b * self.world                      # the FIFO stores the gathered batch; the MLP sees the local one
        ring = model.__dict__.get("_nr_ring")
        if ring is None or not ring.matches(bank0["mb_feat_t"], bank0["mb_feat_v"], x3, mlp_bf16, b):
            ring = BankRing(*[bank0[n] for n in self.bank_names], x3=x3, mlp_bf16=mlp_bf16, batch_rows=b)
            model.__dict__["_nr_ring"] = ring
        else:
            ring.load(*[bank0[n] for n in self.bank_names])
        model.__dict__["_nr_ring_live"] = True
        self._fifo_rows = rows
        return ring

    def set_bank(self, bank):
        """Overwrite the memory bank IN PLACE (the captured graph holds the storage): `bank` has mb_ind, mb_feat_t,
        mb_feat_v, mb_mask_t, mb_mask_v (any device)."""
        if self.ring is not None:
            dev = self.ring.device
            self.ring.load(*[getattr(bank, n).to(dev) for n in self.bank_names])
            self.model.__dict__["_nr_ring_live"] = True
            return
        with torch.no_grad():
            for n in self.bank_names:
                getattr(self.model, n).copy_(getattr(bank, n))

    def _zero_grads(self):
        pass                              # torch.autograd.grad never touches a .grad attribute

    def _body(self):
        m, s = self.model, self.static
        cfg = m.config
        ring = self.ring
        if self.world == 1:
            logit_scale = m.clip.logit_scale.exp()
            # with a ring the mb_* arguments only carry shapes (ring order); the prepared operands come from the ring
            bank = (ring.feat_t, ring.feat_v, ring.mask_t, ring.mask_v) if ring is not None else \
                (m.mb_feat_t, m.mb_feat_v, m.mb_mask_t, m.mb_mask_v)
            losses = m._compute_losses(s["text_feat"], s["video_feat"], s["text_mask"], s["video_mask"], *bank,
                                       cfg.centrality_scale, cfg.beta, cfg.num_neighbors, cfg.temperature, logit_scale,
                                       global_feats=(s["global_text"], s["global_video"]), bank_ring=ring)
            new_rows = (s["idx"], s["text_feat"], s["video_feat"], s["text_mask"], s["video_mask"])
        else:       # row-block sharded head; the NCCL collectives are captured in the graph
            losses, (ta, va, tma, vma, ia) = m._sharded_losses(s["text_feat"], s["video_feat"], s["text_mask"],
                                                               s["video_mask"], (s["global_text"], s["global_video"]),
                                                               idx=s["idx"], bank_ring=ring, defer_text=True)
            new_rows = (ia, ta, va, tma, vma)
        # With bf16 weight-MLP GEMMs the backward reads bf16 copies, never the bank itself: the FIFO update can then
        # leave the critical path and run on its own branch next to the backward.
        early_fifo = (ring is None and self.world == 1 and m._mlp_precision() == "bf16"
                      and os.environ.get("NR_EARLY_FIFO", "1") != "0")
        if early_fifo:
            main = torch.cuda.current_stream()
            self._fifo_stream.wait_stream(main)
            with torch.cuda.stream(self._fifo_stream):
                self._fifo(new_rows)
        out5 = getattr(m, "last_out5", None)
        if out5 is not None and not out5.requires_grad:
            out5 = None
        # torch.autograd.grad hands back the tensors the backward nodes produced (no AccumulateGrad copy that could read
        # a gradient whose collective still runs on a side stream); d total / d out5 = e0 skips the unbind/stack
        # bookkeeping kernels of losses[0].backward()
        leaves = [s[f] for f in GRAD_FIELDS] + self.params
        # split bank insert (single GPU, bf16 weight MLPs on the ring's operand): the backward nodes leave events behind
        # their last reads of the ring
        split_insert = self.world == 1 and ring is not None and ring.mlp_t is not None \
            and os.environ.get("NR_SPLIT_INSERT", "1") != "0"
        if split_insert:
            ops.EVENTS["_want_bank_events"] = True
        if out5 is not None:
            self.grad_list = torch.autograd.grad(out5, leaves, grad_outputs=self._e0, allow_unused=True)
        else:
            self.grad_list = torch.autograd.grad(losses[0], leaves, allow_unused=True)
        m.last_out5 = None
        ops.EVENTS.pop("_want_bank_events", None)
        ev_c, ev_m = ops.EVENTS.pop("contraction_bwd_done", None), ops.EVENTS.pop("mlp_gemm_bwd_done", None)
        if split_insert and ev_c is not None and ev_m is not None:
            # the contraction operands, raw rows, masks and indices of the new samples go in right behind the backward
            # contraction, next to the weight-MLP backward; only the MLP operand waits for the backward GEMM, and that
            # last piece runs next to autograd's final gradient sums instead of after them
            main = torch.cuda.current_stream()
            self._fifo_stream.wait_event(ev_c)
            with torch.cuda.stream(self._fifo_stream), torch.no_grad():
                ring.insert(*new_rows, phase="early")
                self._fifo_stream.wait_event(ev_m)
                ring.insert(*new_rows, phase="late")
            main.wait_stream(self._fifo_stream)
            return out5.detach() if out5 is not None else torch.stack([x.detach() for x in losses])
        if early_fifo:
            main.wait_stream(self._fifo_stream)
        else:
            part = ops.EVENTS.pop("loss_out5_partial", None)
            if part is not None:                # no head parameter took a gradient: reduce the partial losses here
                torch.distributed.all_reduce(part.detach(), op=torch.distributed.ReduceOp.SUM)
            evv = ops.EVENTS.pop("video_grad_ready", None)
            if evv is not None:                 # the video-gradient reduce-scatter ran from a side stream (sharded.py)
                torch.cuda.current_stream().wait_event(evv)
            ev = ops.EVENTS.pop("mlp_backward_done", None) if self.world > 1 else None
            if ring is not None and ev is not None:
                # every reader of the ring is enqueued before this event (it precedes the all-reduce of the head
                # parameters at the end of the backward): the insert runs under that all-reduce on its own stream
                main = torch.cuda.current_stream()
                self._fifo_stream.wait_event(ev)
                with torch.cuda.stream(self._fifo_stream), torch.no_grad():
                    m.wait_gathered_text()
                    ring.insert(*new_rows)
                main.wait_stream(self._fifo_stream)
                return out5.detach() if out5 is not None else torch.stack([x.detach() for x in losses])
            if self.world > 1:
                m.wait_gathered_text()
            if ring is not None:
                # after every reader of the step (both contractions, the MLP GEMMs): only the new rows are written
                with torch.no_grad():
                    ring.insert(*new_rows)
            else:
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
            self._stage_slab, self._stage = _slab_like([self.static[f] for f in FIELDS], dev)
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
            self._static_slab.copy_(self._stage_slab, non_blocking=True)      # all seven inputs: one copy
            self._stage_free.record(cur)
        else:
            for f, t in zip(FIELDS, batch):
                self.static[f].data.copy_(t, non_blocking=True)
        self.graph.replay()
        ops.LAUNCHES["count"] += self.launches_per_replay
        if prefetched and prefetch_next is not None:      # issued AFTER the replay: the host enqueues the next batch's
            self.prefetch(*prefetch_next)                 # copies while the device already runs this step
        if self.ring is not None:
            self.ring.note_replay(self._fifo_rows)
        if sync_losses_to is not None:
            sync_losses_to.copy_(self.losses, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return sync_losses_to
        return self.losses


# ---- trainer-compatible wrapper: model(...) returns losses whose .backward() hands out the captured gradients ----
class _ReplayFunction(torch.autograd.Function):
    """forward: copy the batch into the static buffers and replay the captured forward + backward + bank FIFO;
    backward: d total * (the gradients the replay already produced).  Only `total` carries a gradient — the four
    component losses are returned detached (the reference's trainer only logs them, training/trainer.py:84-125)."""

    @staticmethod
    def forward(ctx, runner, text_mask, video_mask, idx, text, video, gt, gv, *params):
        step = runner.step
        with torch.no_grad():
            for f, t in zip(FIELDS, (text, video, text_mask, video_mask, idx, gt, gv)):
                step.static[f].data.copy_(t.reshape(step.static[f].shape), non_blocking=True)
        step.graph.replay()
        ops.LAUNCHES["count"] += step.launches_per_replay
        if step.ring is not None:
            step.ring.note_replay(step._fifo_rows)
        ctx.runner = runner
        ctx.token = runner.replays = runner.replays + 1
        ctx.meta = [(t.dtype, t.shape) for t in (text, video, gt, gv)]
        return step.losses.clone()

    @staticmethod
    def backward(ctx, g5):
        r = ctx.runner
        if ctx.token != r.replays:
            raise RuntimeError("GraphedHead: backward() of a step whose static gradients were overwritten by a later "
                               "forward; call loss.backward() before the next model(...) (or set head_graph=False)")
        s = g5[0]
        gl = r.step.grad_list
        live = [g for g in gl if g is not None]
        scaled = iter(torch._foreach_mul(live, s))              # one multi-tensor launch for all 13 gradients
        out = []
        for i, g in enumerate(gl):
            if g is None:
                out.append(None)
                continue
            t = next(scaled)
            if i < 4:
                dt, shp = ctx.meta[i]
                t = t.to(dt).reshape(shp)
            out.append(t)
        return (None, None, None, None, *out)


class GraphedHead:
    """Per-model cache of captured head steps keyed by the batch / bank shapes (HeadMixin.head_forward uses it when
    `head_graph` is on).  The bank lives in static storage owned by the captured graph: tensors assigned to
    `model.mb_*` from outside (MemoryBankManager.load_memory_bank, reference utils/memory_bank.py:206-211) are
    detected by identity, copied into the static storage and re-pointed at it."""
    MAX_SHAPES = 4

    def __init__(self, model):
        self.runners = {}
        self.model = model

    def _key(self, text, video, gt, gv, tm, vm):
        m = self.model
        return (tuple(text.shape), tuple(video.shape), tuple(gt.shape), tuple(gv.shape), tuple(tm.shape), tuple(vm.shape),
                tuple(m.mb_feat_t.shape), tuple(m.mb_feat_v.shape), m.mb_feat_t.dtype, m.mb_mask_t.dtype,
                m.mb_ind.dtype, getattr(m.config, "world_size", 1), m._head_precision(), m._head_bwd_precision(),
                m._mlp_precision(), text.device.index)

    def __call__(self, text, video, tm, vm, idx, gt, gv):
        m = self.model
        key = self._key(text, video, gt, gv, tm, vm)
        r = self.runners.get(key)
        if r is not None and r.step.ring is not None and m.__dict__.get("_nr_ring") is not r.step.ring:
            r = None                                          # another capture replaced the model's bank ring
        if r is None:
            if len(self.runners) >= self.MAX_SHAPES:
                return None                                   # too many distinct shapes: the caller runs eagerly
            r = self.runners[key] = _Runner(m, (text, video, tm, vm, idx, gt, gv))
        r.sync_bank(m)
        out5 = _ReplayFunction.apply(r, tm, vm, idx, text, video, gt, gv, *r.step.params)
        m.last_neighbors = r.neighbors
        total = out5[0]
        c, u, n, k = out5.detach()[1:].unbind(0)
        return total, c, u, n, k


class _Runner:
    def __init__(self, model, example):
        bank_in = {n: getattr(model, n) for n in ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v")}
        self.step = GraphedHeadStep(model, [t.detach() for t in example], explicit_grads=True)
        self.neighbors = getattr(model, "last_neighbors", None)
        self.bank = {n: getattr(model, n) for n in self.step.bank_names}     # the static storage
        self.replays = 0
        del bank_in

    def sync_bank(self, model):
        """Static bank storage <- whatever was assigned to model.mb_* since the last step (identity check; with a
        ring: the `_nr_ring_live` flag the attribute setters clear)."""
        ring = self.step.ring
        if ring is not None:
            if model.__dict__.get("_nr_ring") is not ring:
                raise RuntimeError("GraphedHead: the model's bank ring was replaced; re-capture the step")
            if not model.__dict__.get("_nr_ring_live"):
                ring.load(*[getattr(model, n) for n in self.step.bank_names])
                model.__dict__["_nr_ring_live"] = True
            return
        for n, st in self.bank.items():
            cur = getattr(model, n)
            if cur is not st:
                with torch.no_grad():
                    st.copy_(cur)
                setattr(model, n, st)
