"""The whole retrieval head below the token-weight MLPs as ONE autograd node.

`_compute_losses` of the reference (NeighborRetr/models/modeling.py:314-360) expands into ~400 ATen launches per
step; at b=128 the head is launch-bound, so the fused node issues the ~60 C-ABI launches of a step back to back
with no autograd bookkeeping in between, which also makes the step capturable in a CUDA graph (graph.py).

Data flow (W = 1: the row block is the full batch; r0/rows are already threaded through for row sharding):
  prep(text, video, bank_t, bank_v)                      -> normalised fp32 + bf16 operand copies, column sums
  S, S^T           = 1/2 (H(text,video) + H(video,text)^T)                       2 max-sim launches
  mb_t2v, mb_v2t   = bank similarities [B,M]                                      4 max-sim launches
  c_t2v, c_v2t     = row means (bank centrality, until_module.py:181)
  G, G^T           = gT gV^T (library GEMM), Sinkhorn duals for both directions   1 cluster launch
  w_t, w_v         = centrality weights (column mean + GEMV)
  row losses       = both directions, all four losses in one pass per row         2 launches
Backward mirrors it: row backward -> dS, dG, dc, dw, dls; tensor-core routing products for the batch pair and
the four bank pairs (the bank gradient dH is the rank-1 broadcast dc[b]/M, passed as a stride-0 view).
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import (NR_LOSS_CENTRALITY, NR_LOSS_KL, NR_LOSS_NEIGHBOR, NR_LOSS_UNIFORM, NR_NSAVE, NR_PREC_BF16,
                   NR_PREC_BF16X3)
from .ops import Prepared, _call, _f32c, _mask, _p, _req_cuda, _stream

ALL_LOSSES = NR_LOSS_CENTRALITY | NR_LOSS_NEIGHBOR | NR_LOSS_KL | NR_LOSS_UNIFORM


def _fwd_dir(prec, X, Y, wx, mx, my, out, sr, sc, out2, sr2, sc2, acc):
    return ops._maxsim_dir_fwd(prec, X, Y, wx, mx, my, 0.5, out, sr, sc, out2, sr2, sc2, acc, True)


_M54 = {}


def _combine_matrix(B, wu, wn, wkl, dev):
    """[total, centrality, uniform, neighbor, kl] = M54 @ raw row sums {centrality, neighbour, kl, uniform}
    (means over B rows — B^2 entries for KL — and over the two directions; reference modeling.py:353-358).
    Cached: creating it costs a host->device copy, which must not happen inside a CUDA-graph capture."""
    key = (B, wu, wn, wkl, str(dev))
    m = _M54.get(key)
    if m is None:
        c1, c2 = 0.5 / B, 0.5 / (B * B)
        m = torch.tensor([[c1, wn * c1, wkl * c2, wu * c1], [c1, 0, 0, 0], [0, 0, 0, c1], [0, c1, 0, 0],
                          [0, 0, c2, 0]], dtype=torch.float32, device=dev)
        _M54[key] = m
    return m


class HeadPrologue:
    """Everything of a head step that does not depend on the token-weight MLPs: token preparation (x4, with the
    transposed operand copies of the backward), centrality weights, global similarity and the Sinkhorn duals.
    The constructor only ALLOCATES (on the current = main stream); the three run_* groups are independent and may
    be enqueued on forked streams next to the MLP evaluations (modeling._compute_losses)."""

    def __init__(self, text, video, gt, gv, text_mask, video_mask, mb_feat_t, mb_feat_v, mb_mask_t, mb_mask_v, hp,
                 need_text_grad=True, need_video_grad=True, bank_ring=None):
        """bank_ring (bank.BankRing): the bank's prepared operands already exist and persist across steps; nothing of
        the bank is prepared here (mb_* are then only read for their shapes)."""
        cs, beta, k, tau, iters, wu, wn, wkl, prec, bprec = hp
        _req_cuda(text, video, gt, gv, mb_feat_t, mb_feat_v)
        dev = text.device
        self.hp = hp
        self.tm, self.vm = _mask(text_mask), _mask(video_mask)
        self.bank_static = bank_ring is not None
        self.mtm, self.mvm = (bank_ring.mask_t, bank_ring.mask_v) if self.bank_static else (_mask(mb_mask_t), _mask(mb_mask_v))
        self.bf = bf = prec in ops.TC_PRECISIONS or bprec in ops.TC_PRECISIONS
        x3 = prec == NR_PREC_BF16X3
        # bf16: masks are folded into the operand copies (masked tokens = zero rows) for the two-direction kernel
        self.fusedk = fk = (prec in ops.TC_PRECISIONS and bprec == prec and ops.USE_FUSED_MAXSIM
                            and ops.maxsim2_supported(text.shape[1], video.shape[1], text.shape[2] * (3 if x3 else 1)))
        if x3 and not fk:
            raise RuntimeError("precision 'bf16x3' needs the fused two-direction kernel for these token counts")
        rx, ry = (ops.ROLE_X, ops.ROLE_Y) if x3 else (0, 0)      # the text side is X, the video side Y (HeadFunction)
        self.T = Prepared(text.detach(), bf16=bf, colsum=True, mask=self.tm if fk else None, defer=True, split=rx)
        self.V = Prepared(video.detach(), bf16=bf, colsum=True, mask=self.vm if fk else None, defer=True, split=ry)
        if self.bank_static:
            if not fk or bank_ring.x3 != x3:
                raise RuntimeError("bank ring: needs the fused tensor-core path in the precision it was built for")
            self.MT, self.MV = bank_ring.MT, bank_ring.MV
        else:
            self.MT = Prepared(mb_feat_t, bf16=bf, mask=self.mtm if fk else None, defer=True, f32=not fk, split=rx)
            self.MV = Prepared(mb_feat_v, bf16=bf, mask=self.mvm if fk else None, defer=True, f32=not fk, split=ry)
        B, M, d = self.T.r, self.MT.r, self.T.d
        if self.V.r != B or self.MV.r != M:
            raise RuntimeError("text/video batch sizes (or bank sizes) differ")
        f32 = dict(dtype=torch.float32, device=dev)
        self.g2, self.v2 = _f32c(gt.detach()).reshape(B, d), _f32c(gv.detach()).reshape(B, d)
        self.GG = torch.empty(2, B, B, **f32)              # [G ; G^T]
        self.duals = torch.empty(4, B, **f32)
        self.lib_ws, self.nws = ops.sinkhorn_workspace(B, dev)
        self.mean = torch.empty(2, d, **f32); self.gn = torch.empty(2, B, d, **f32)
        self.ginv = torch.empty(2, B, **f32); self.w = torch.empty(2, B, **f32)
        self.need_t, self.need_v = need_text_grad, need_video_grad
        self.global_done = None            # event of a still-running run_global() branch (see modeling.py)
        if bf:
            for P in (self.T, self.V, self.MT, self.MV):
                P.alloc_transposed()

    def _centrality(self, P, g, i):
        B, d, cs = self.T.r, self.T.d, self.hp[0]
        _call("nr_centrality_fwd", _p(P.partials), P.partials.shape[0], P.rows, _p(g), B, d, cs, _p(self.mean[i]),
              _p(self.gn[i]), _p(self.ginv[i]), _p(self.w[i]), _stream(), launches=2)

    def run_text_side(self):
        bprec = self.hp[9]
        if not self.bank_static:
            self.MT.run()
        self.T.run()
        self._centrality(self.T, self.g2, 0)
        if self.bf and self.need_v:        # sources of the video-side backward contraction, off its critical path
            self.T.bwd_source(bprec)
            if not self.bank_static:
                self.MT.bwd_source(bprec)

    def run_video_side(self):
        bprec = self.hp[9]
        if not self.bank_static:
            self.MV.run()
        self.V.run()
        self._centrality(self.V, self.v2, 1)
        if self.bf and self.need_t:
            self.V.bwd_source(bprec)
            if not self.bank_static:
                self.MV.bwd_source(bprec)

    def run_global(self):
        # global similarity: one token per sample -> plain dot products (library GEMM, fp32), then the Sinkhorn duals
        B, iters = self.T.r, int(self.hp[4])
        _call("nr_gram_f32", _p(self.g2), _p(self.v2), B, B, self.T.d, _p(self.GG[0]), _p(self.GG[1]), _stream())
        d_ = self.duals
        # in-step: 16 rows per CTA for the multi-CTA variants (the chain shares the GPU with the contraction)
        _call("nr_sinkhorn_ex", _p(self.GG[0]), _p(self.GG[1]), B, iters, _p(d_[0]), _p(d_[1]), _p(d_[2]), _p(d_[3]),
              _p(self.lib_ws), self.nws, 16 if B >= 256 else 0, _stream())

    def run_forked(self):
        with ops.ForkJoin(2) as fj:
            self.run_text_side()
            with fj.on(0):
                self.run_video_side()
            with fj.on(1):
                self.run_global()
        return self


class HeadFunction(torch.autograd.Function):
    """(text, video, gT, gV, tw, vw, tw_mb, vw_mb, logit_scale) -> [total, centrality, uniform, neighbor, kl].

    Independent launch groups run on forked streams (ops.ForkJoin): inside a CUDA graph they become parallel
    branches, which matters at b=128 where most kernels fill a fraction of the GPU."""

    @staticmethod
    def forward(ctx, text, video, gt, gv, tw, vw, tw_mb, vw_mb, logit_scale, text_mask, video_mask, mb_feat_t,
                mb_feat_v, mb_mask_t, mb_mask_v, hp, pro=None):
        cs, beta, k, tau, iters, wu, wn, wkl, prec, bprec = hp
        _req_cuda(text, video, gt, gv, tw, vw, tw_mb, vw_mb, logit_scale, mb_feat_t, mb_feat_v)
        ctx.set_materialize_grads(False)         # no zero-filled gradient for the neighbour lists
        dev = text.device
        tw, vw, tw_mb, vw_mb = _f32c(tw), _f32c(vw), _f32c(tw_mb), _f32c(vw_mb)
        if pro is None:                 # stand-alone use: the prologue runs here (already-joined when passed in)
            need = ctx.needs_input_grad
            pro = HeadPrologue(text, video, gt, gv, text_mask, video_mask, mb_feat_t, mb_feat_v, mb_mask_t,
                               mb_mask_v, hp, need[0], need[1]).run_forked()
        T, V, MT, MV = pro.T, pro.V, pro.MT, pro.MV
        tm, vm, mtm, mvm = pro.tm, pro.vm, pro.mtm, pro.mvm
        fusedk = pro.fusedk
        B, M, d = T.r, MT.r, T.d
        g2, v2, G, GT, duals = pro.g2, pro.v2, pro.GG[0], pro.GG[1], pro.duals
        mean, gn, ginv, w = pro.mean, pro.gn, pro.ginv, pro.w
        f32 = dict(dtype=torch.float32, device=dev)
        S = torch.empty(B, B, **f32); ST = torch.empty(B, B, **f32)
        mb = torch.empty(2, B, M, **f32)                  # [mb_t2v ; mb_v2t]
        mb_t2v, mb_v2t = mb[0], mb[1]
        cb = torch.empty(2, B, **f32)                     # [c_t2v ; c_v2t]
        ls = _f32c(logit_scale).reshape(1)
        row_out = torch.empty(2, 4, B, **f32)             # ALL_LOSSES: every entry is written by the row kernels
        nbr = torch.empty(2, B, k, dtype=torch.int32, device=dev)
        saved = torch.empty(2, B, NR_NSAVE, **f32)
        sums = torch.empty(8, **f32)
        m54 = _combine_matrix(B, wu, wn, wkl, dev)
        st = _stream()
        if fusedk:
            # ONE launch: the batch pair and both bank pairs, each token pair multiplied once (the larger problems
            # first so that the tail of the persistent tile list is the small one)
            svA, svC, sv1 = ops.maxsim2_fwd([
                dict(X=T, Y=MV, wx=tw, wy=vw_mb, alpha=0.5, out=mb_t2v, strides=(M, 1)),      # S(text, bank_v)
                dict(X=MT, Y=V, wx=tw_mb, wy=vw, alpha=0.5, out=mb_v2t, strides=(1, M)),      # S(bank_t, video)^T
                dict(X=T, Y=V, wx=tw, wy=vw, alpha=0.5, out=S, strides=(B, 1), out2=ST, strides2=(1, B))])
            p1, y1, p2, y2 = sv1          # pmax_x, ystar, pmax_y, xstar of the batch pair
            pA, yA, pB, yB = svA
            pC, yC, pD, yD = svC
        else:
            p1, y1 = _fwd_dir(prec, T, V, tw, tm, vm, S, B, 1, ST, 1, B, 0)
            p2, y2 = _fwd_dir(prec, V, T, vw, vm, tm, S, 1, B, ST, B, 1, 1)
            pA, yA = _fwd_dir(prec, T, MV, tw, tm, mvm, mb_t2v, M, 1, None, 0, 0, 0)        # H(text, bank_v)
            pB, yB = _fwd_dir(prec, MV, T, vw_mb, mvm, tm, mb_t2v, 1, M, None, 0, 0, 1)     # H(bank_v, text)^T
            pD, yD = _fwd_dir(prec, V, MT, vw, vm, mtm, mb_v2t, M, 1, None, 0, 0, 0)        # H(video, bank_t)
            pC, yC = _fwd_dir(prec, MT, V, tw_mb, mtm, vm, mb_v2t, 1, M, None, 0, 0, 1)     # H(bank_t, video)^T
        ctx.fusedk = fusedk
        _call("nr_row_mean", _p(mb), M, 2 * B, M, _p(cb), st)                # both bank centralities at once
        if pro.global_done is not None:                  # G, G^T and the Sinkhorn duals from the detached branch
            torch.cuda.current_stream().wait_event(pro.global_done)
            pro.global_done = None
        # ---- fork 2: row losses, the two directions side by side
        with ops.ForkJoin(1) as fj:
            _call("nr_row_losses_fwd", _p(S), B, _p(G), B, _p(cb[1]), _p(w[0]), _p(duals[0]), _p(duals[1]), B, B, 0,
                  _p(ls), k, tau, tau, beta, ALL_LOSSES, _p(row_out[0]), _p(nbr[0]), _p(saved[0]), _stream())
            with fj.on(0):
                _call("nr_row_losses_fwd", _p(ST), B, _p(GT), B, _p(cb[0]), _p(w[1]), _p(duals[2]), _p(duals[3]), B, B,
                      0, _p(ls), k, tau, tau, beta, ALL_LOSSES, _p(row_out[1]), _p(nbr[1]), _p(saved[1]), _stream())
        # [total, centrality, uniform, neighbor, kl] = M54 @ (sums_dir1 + sums_dir2);  sums order: c, n, kl, u
        out5 = torch.empty(5, **f32)
        ctx.ev_out5 = None
        if torch.cuda.is_current_stream_capturing() and any(ctx.needs_input_grad):
            # captured step: nothing of the backward reads the loss VALUES, so their two small reductions leave the
            # critical path between the row losses and their backward (joined at the end of backward())
            with ops.ForkJoin(1, offset=9) as fj:
                with fj.on(0):
                    _call("nr_vec_sums", _p(row_out), 8, B, None, _p(sums), _stream())
                    _call("nr_matvec_small", _p(m54), 5, 4, 0, _p(sums), _p(sums[4:]), _p(out5), _stream())
                ctx.ev_out5 = fj.detach(0)
            # the branch reads these after forward() has returned: keep them out of the allocator until it is joined
            ctx.keep_out5 = (row_out, sums)
        else:
            _call("nr_vec_sums", _p(row_out), 8, B, None, _p(sums), st)
            _call("nr_matvec_small", _p(m54), 5, 4, 0, _p(sums), _p(sums[4:]), _p(out5), st)
        ctx.hp = hp
        ctx.objs = (T, V, MT, MV)
        # the masks are only read by the one-direction backward kernels (the fused path folded them into the operand
        # copies); not saving them also keeps the bank masks free for an in-place FIFO update before the backward
        sm = (None, None, None, None) if fusedk else (tm, vm, mtm, mvm)
        ctx.save_for_backward(tw, vw, tw_mb, vw_mb, *sm, S, ST, G, GT, cb, duals, w, ls, nbr, saved, mean,
                              gn, ginv, g2, v2, m54, p1, y1, p2, y2, pA, yA, pB, yB, pC, yC, pD, yD)
        ctx.gshape = (gt.shape, gv.shape)
        # the backward's accumulators (token gradients, dc, dw, weight gradients, dls), zeroed next to the forward
        nw_ = tw.numel() + vw.numel() + tw_mb.numel() + vw_mb.numel()
        ctx.prezero = ops.prezeroed((T.rows + V.rows) * d + 4 * B + nw_ + 1, dev, 8) if any(ctx.needs_input_grad) else None
        ctx.nbr = nbr
        ctx.mark_non_differentiable(nbr)
        return out5, nbr

    @staticmethod
    def backward(ctx, g5, _gn):
        (tw, vw, tw_mb, vw_mb, tm, vm, mtm, mvm, S, ST, G, GT, cb, duals, w, ls, nbr, saved, mean, gn, ginv, g2, v2, m54,
         p1, y1, p2, y2, pA, yA, pB, yB, pC, yC, pD, yD) = ctx.saved_tensors
        cs, beta, k, tau, iters, wu, wn, wkl, prec, bprec = ctx.hp
        T, V, MT, MV = ctx.objs
        B, M, d = T.r, MT.r, T.d
        dev = S.device
        need = ctx.needs_input_grad
        nt, nv = T.n, V.n
        f32 = dict(dtype=torch.float32, device=dev)
        # ---- every buffer of the backward, allocated on the main stream before any fork
        gscale = torch.empty(4, **f32)                                # upstream multipliers of the raw row terms
        _call("nr_matvec_small", _p(m54), 5, 4, 1, _p(_f32c(g5)), None, _p(gscale), _stream())
        # one zero-fill: [dtn | dvn | dc_t2v | dc_v2t | dw_t | dw_v | dtw | dvw | dtw_mb | dvw_mb | dls]
        # (the token gradients first: their red.global.add.v4 needs 16-byte alignment)
        nw = tw.numel() + vw.numel() + tw_mb.numel() + vw_mb.numel()
        ntok = (T.rows + V.rows) * d
        if ctx.prezero is not None:
            z, ev_z = ctx.prezero
            torch.cuda.current_stream().wait_event(ev_z)
            ctx.prezero = None
        else:
            z = torch.zeros(ntok + 4 * B + nw + 1, **f32)
        dtn, dvn = z[:T.rows * d], z[T.rows * d:ntok]
        dc, dw = z[ntok:ntok + 2 * B].view(2, B), z[ntok + 2 * B:ntok + 4 * B].view(2, B)
        o = ntok + 4 * B
        dtw = z[o:o + tw.numel()]; o += tw.numel()
        dvw = z[o:o + vw.numel()]; o += vw.numel()
        dtw_mb = z[o:o + tw_mb.numel()]; o += tw_mb.numel()
        dvw_mb = z[o:o + vw_mb.numel()]; o += vw_mb.numel()
        dls = z[o:o + 1]
        dS1 = torch.empty(B, B, **f32); dS2 = torch.empty(B, B, **f32)
        dG1 = torch.empty(B, B, **f32); dG2 = torch.empty(B, B, **f32)
        dS = torch.empty(B, B, **f32); dG = torch.empty(B, B, **f32)
        dgt = torch.empty(B, d, **f32) if need[2] else None
        dgv = torch.empty(B, d, **f32) if need[3] else None
        dmean = torch.empty(2, d, **f32)
        dtext = torch.empty_like(T.xn) if need[0] else None
        dvideo = torch.empty_like(V.xn) if need[1] else None
        # ---- fork 1: row-loss backward, both directions
        with ops.ForkJoin(1) as fj:
            _call("nr_row_losses_bwd", _p(S), B, _p(G), B, _p(cb[1]), _p(w[0]), _p(duals[0]), _p(duals[1]), B, B, 0,
                  _p(ls), k, tau, tau, beta, ALL_LOSSES, _p(nbr[0]), _p(saved[0]), _p(gscale), _p(dS1), B, _p(dG1), B,
                  _p(dc[1]), _p(dw[0]), _p(dls), _stream())
            with fj.on(0):
                _call("nr_row_losses_bwd", _p(ST), B, _p(GT), B, _p(cb[0]), _p(w[1]), _p(duals[2]), _p(duals[3]), B, B,
                      0, _p(ls), k, tau, tau, beta, ALL_LOSSES, _p(nbr[1]), _p(saved[1]), _p(gscale), _p(dS2), B,
                      _p(dG2), B, _p(dc[0]), _p(dw[1]), _p(dls), _stream())
        # ---- fork 2: token-pair contractions (main: text side, side 0: video side), global path (side 1),
        #      token-weight gradients (side 2)
        # (the global branches on high-priority streams: their small kernels are placed as soon as an SM has room
        # instead of queueing behind the ~1700 pending CTAs of the weight-gradient reduction)
        with ops.ForkJoin(4, high=(1, 3)) as fj:
            _call("nr_transpose_add", _p(dS1), B, _p(dS2), B, _p(dS), B, B, B, 1.0, 1.0, _stream())
            ev_dS = torch.cuda.Event()
            ev_dS.record()

            def global_path(ev_dG=None):
                # centrality backward FIRST (it only needs dw from the row losses and writes d mean, which the token
                # normalisation backward at the end of the critical path is waiting for); then dgT += dG gV and
                # dgV += dG^T gT accumulate on top of it, off the critical path.  The video half runs on its own branch.
                if ev_dG is None:
                    if need[2]:
                        _call("nr_centrality_bwd", _p(mean[0]), _p(gn[0]), _p(ginv[0]), _p(w[0]), _p(dw[0]), B, d, cs,
                              T.rows, _p(dgt), 0, _p(dmean[0]), _stream())
                    else:
                        _call("nr_centrality_bwd", _p(mean[0]), _p(gn[0]), _p(ginv[0]), _p(w[0]), _p(dw[0]), B, d, cs,
                              T.rows, None, 0, _p(dmean[0]), _stream())
                    ev_dm.append(torch.cuda.Event()); ev_dm[-1].record()
                    _call("nr_transpose_add", _p(dG1), B, _p(dG2), B, _p(dG), B, B, B, 1.0, 1.0, _stream())
                    ev = torch.cuda.Event()
                    ev.record()
                    if need[2]:
                        _call("nr_matmul_f32", _p(dG), B, 0, _p(v2), d, B, B, d, _p(dgt), d, 1, _stream())
                    return ev
                _call("nr_centrality_bwd", _p(mean[1]), _p(gn[1]), _p(ginv[1]), _p(w[1]), _p(dw[1]), B, d, cs, V.rows,
                      _p(dgv) if need[3] else None, 0, _p(dmean[1]), _stream())
                ev_dm.append(torch.cuda.Event()); ev_dm[-1].record()
                torch.cuda.current_stream().wait_event(ev_dG)
                if need[3]:
                    _call("nr_matmul_f32", _p(dG), B, 1, _p(g2), d, B, B, d, _p(dgv), d, 1, _stream())
                return None

            ev_dm = []
            with fj.on(1):
                ev_dG = global_path()
            with fj.on(3):
                global_path(ev_dG)
            # the two branches stay open past this fork: the normalisation backward only waits for their d mean
            # events, the dG products are joined right before the gradients are handed back
            ev_global_end = (fj.detach(1), fj.detach(3))
            if ctx.fusedk:
                # one routing matrix per pair, applied from either side; ALL contractions in one launch (the text
                # gradient accumulates over [video ; bank-video] sources, the video gradient over [text ; bank-text])
                sc = 0.5 / M
                jobs = []
                if need[0]:
                    jobs += [(0, V, tw, vw, y1, y2, dS, B, 1, 0.5, B, B, dtn),
                             (0, MV, tw, vw_mb, yA, yB, dc[0], 1, 0, sc, B, M, dtn)]
                if need[1]:
                    jobs += [(1, T, tw, vw, y1, y2, dS, B, 1, 0.5, B, B, dvn),
                             (1, MT, tw_mb, vw, yC, yD, dc[1], 0, 1, sc, M, B, dvn)]
                if jobs:
                    ops.maxsim2_bwd_multi(jobs, nt, nv, d)
                if ops.EVENTS.get("_want_bank_events"):    # last reader of the bank's contraction operands (graph.py)
                    ev_c = torch.cuda.Event()
                    ev_c.record()
                    ops.EVENTS["contraction_bwd_done"] = ev_c
                with fj.on(2):
                    torch.cuda.current_stream().wait_event(ev_dS)
                    ops.maxsim2_bwd_w_multi([
                        (p1, p2, dS, B, 1, 0.5, B, B, dtw if need[4] else None, dvw if need[5] else None),
                        (pA, pB, dc[0], 1, 0, sc, B, M, dtw if need[4] else None, dvw_mb if need[7] else None),
                        (pC, pD, dc[1], 0, 1, sc, M, B, dtw_mb if need[6] else None, dvw if need[5] else None)], nt, nv)
            else:
                st = _stream()
                vs, vld = V.bwd_source(bprec); ts, tld = T.bwd_source(bprec)
                mvs, mvld = MV.bwd_source(bprec); mts, mtld = MT.bwd_source(bprec)
                if need[0]:
                    # text <- batch pair (both orientations) and the text-vs-bank-video pair
                    _call("nr_maxsim_bwd_x", bprec, _p(vs), vld, _p(tw), _p(tm), _p(vm), _p(y1), _p(dS), B, 1, 0.5, B, nt,
                          B, nv, d, _p(dtn), st)
                    _call("nr_maxsim_bwd_y", bprec, _p(vs), vld, _p(vw), _p(vm), _p(tm), _p(y2), _p(dS), 1, B, 0.5, B, nv,
                          B, nt, d, _p(dtn), st)
                    _call("nr_maxsim_bwd_x", bprec, _p(mvs), mvld, _p(tw), _p(tm), _p(mvm), _p(yA), _p(dc[0]), 1, 0,
                          0.5 / M, B, nt, M, nv, d, _p(dtn), st)
                    _call("nr_maxsim_bwd_y", bprec, _p(mvs), mvld, _p(vw_mb), _p(mvm), _p(tm), _p(yB), _p(dc[0]), 0, 1,
                          0.5 / M, M, nv, B, nt, d, _p(dtn), st)
                if need[1]:
                    _call("nr_maxsim_bwd_y", bprec, _p(ts), tld, _p(tw), _p(tm), _p(vm), _p(y1), _p(dS), B, 1, 0.5, B, nt,
                          B, nv, d, _p(dvn), st)
                    _call("nr_maxsim_bwd_x", bprec, _p(ts), tld, _p(vw), _p(vm), _p(tm), _p(y2), _p(dS), 1, B, 0.5, B, nv,
                          B, nt, d, _p(dvn), st)
                    _call("nr_maxsim_bwd_x", bprec, _p(mts), mtld, _p(vw), _p(vm), _p(mtm), _p(yD), _p(dc[1]), 1, 0,
                          0.5 / M, B, nv, M, nt, d, _p(dvn), st)
                    _call("nr_maxsim_bwd_y", bprec, _p(mts), mtld, _p(tw_mb), _p(mtm), _p(vm), _p(yC), _p(dc[1]), 0, 1,
                          0.5 / M, M, nt, B, nv, d, _p(dvn), st)
                # token-weight gradients (batch pair + bank pairs)
                if need[4]:
                    _call("nr_maxsim_bwd_w", _p(p1), _p(dS), B, 1, 0.5, B, nt, B, _p(dtw), st)
                    _call("nr_maxsim_bwd_w", _p(pA), _p(dc[0]), 1, 0, 0.5 / M, B, nt, M, _p(dtw), st)
                if need[5]:
                    _call("nr_maxsim_bwd_w", _p(p2), _p(dS), 1, B, 0.5, B, nv, B, _p(dvw), st)
                    _call("nr_maxsim_bwd_w", _p(pD), _p(dc[1]), 1, 0, 0.5 / M, B, nv, M, _p(dvw), st)
                if need[6]:
                    _call("nr_maxsim_bwd_w", _p(pC), _p(dc[1]), 0, 1, 0.5 / M, M, nt, B, _p(dtw_mb), st)
                if need[7]:
                    _call("nr_maxsim_bwd_w", _p(pB), _p(dc[0]), 0, 1, 0.5 / M, M, nv, B, _p(dvw_mb), st)
            # the weight-gradient reduction outlasts the contraction: the normalisation backward does not wait for it
            ev_global_end = ev_global_end + (fj.detach(2),)
        # ---- fork 3: normalisation backward of the two modalities
        for ev in ev_dm:
            torch.cuda.current_stream().wait_event(ev)
        with ops.ForkJoin(1) as fj:
            if need[0]:
                T.backward(dtn, add_vec=dmean[0], out=dtext)
            with fj.on(0):
                if need[1]:
                    V.backward(dvn, add_vec=dmean[1], out=dvideo)
        for ev in ev_global_end:
            torch.cuda.current_stream().wait_event(ev)
        if ctx.ev_out5 is not None:
            torch.cuda.current_stream().wait_event(ctx.ev_out5)
            ctx.keep_out5 = None
        ctx.objs = None
        gs_t, gs_v = ctx.gshape
        return (dtext, dvideo, dgt.reshape(gs_t) if need[2] else None, dgv.reshape(gs_v) if need[3] else None,
                dtw.view_as(tw) if need[4] else None, dvw.view_as(vw) if need[5] else None,
                dtw_mb.view_as(tw_mb) if need[6] else None, dvw_mb.view_as(vw_mb) if need[7] else None,
                dls.reshape(()) if need[8] else None, None, None, None, None, None, None, None, None)


def fused_head(text, video, gt, gv, tw, vw, tw_mb, vw_mb, logit_scale, text_mask, video_mask, mb_feat_t, mb_feat_v,
               mb_mask_t, mb_mask_v, *, centrality_scale, beta, num_neighbors, temperature, uniform_weight,
               neighbor_weight, kl_weight, precision="bf16", bwd_precision=None, iters=50, prologue=None):
    hp = head_hparams(centrality_scale, beta, num_neighbors, temperature, uniform_weight, neighbor_weight, kl_weight,
                      precision, bwd_precision, iters)
    return HeadFunction.apply(text, video, gt, gv, tw, vw, tw_mb, vw_mb, logit_scale, text_mask, video_mask,
                              mb_feat_t, mb_feat_v, mb_mask_t, mb_mask_v, hp, prologue)


def head_hparams(centrality_scale, beta, num_neighbors, temperature, uniform_weight, neighbor_weight, kl_weight,
                 precision="bf16", bwd_precision=None, iters=50):
    prec = ops.PRECISIONS[precision]
    bprec = prec if prec == NR_PREC_BF16X3 else ops.PRECISIONS[bwd_precision or precision]
    return (float(centrality_scale), float(beta), int(num_neighbors), float(temperature), int(iters),
            float(uniform_weight), float(neighbor_weight), float(kl_weight), prec, bprec)
