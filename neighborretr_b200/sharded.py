"""Row-block sharded retrieval head for W > 1 ranks.

The reference replicates the whole [B,B] head on every rank after the gather (modeling.py:274-312).  Here rank r
owns rows r*b.. of S (text->video direction) and of S^T (video->text direction): two [b,B] blocks plus its rows of
the two bank similarities, so all per-row work (top-k, min/max, log-sum-exps, KL, InfoNCE, neighbour loss) is
rank-local.  Exchanges: all_gather of the (features, masks, global features, token weights) and of the bank
centrality vectors c [b]->[B]; all_reduce of the 8 loss partial sums; in backward all_reduce of dc / d logit_scale
and reduce_scatter(SUM) of the gradients w.r.t. the gathered tensors.  Every rank returns the GLOBAL loss values;
its feature gradients equal rows [r*b,(r+1)*b) of the single-process full-batch gradient (what AllGather.backward
yields in the reference), and head-parameter gradients equal the full gradients once summed over ranks
(`SumGradAcrossRanks`), as in the reference where every rank holds the full head gradient.

Sinkhorn and the global similarity G [B,B] are replicated (2*B^2*D flops, no exchange inside the 100 iterations).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import ops
from ._lib import NR_NSAVE, NR_PREC_BF16, NR_PREC_BF16X3
from .fused import ALL_LOSSES, _combine_matrix, _fwd_dir, head_hparams
from .ops import Prepared, _call, _f32c, _mask, _p, _req_cuda, _stream


# ---- collectives (contiguous buffers, default process group) ------------------------------------------------------
def _gather(t):
    t = t.contiguous()
    w = dist.get_world_size()
    out = torch.empty((w * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return out


def _reduce_scatter(t, b):
    """Sum over ranks of t [W*b, ...] and return this rank's rows [b, ...]."""
    t = t.contiguous()
    out = torch.empty((b,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    if dist.get_backend() == "nccl":
        dist.reduce_scatter_tensor(out, t, op=dist.ReduceOp.SUM)
    else:                                   # gloo on CUDA tensors (single-GPU emulation in tests)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        r = dist.get_rank()
        out.copy_(t[r * b:(r + 1) * b])
    return out


def _all_to_all_blocks(send):
    """send [W, ...]: chunk q goes to rank q; returns recv [W, ...] whose chunk r came from rank r."""
    send = send.contiguous()
    recv = torch.empty_like(send)
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(recv, send)
    else:                                   # gloo on CUDA tensors (single-GPU emulation in tests)
        W, r = dist.get_world_size(), dist.get_rank()
        allb = torch.empty((W * send.shape[0],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(allb, send)
        recv.copy_(allb.view((W,) + tuple(send.shape))[:, r])
    return recv


class SumGradsAcrossRanks(torch.autograd.Function):
    """Identity on a group of parameters whose backward all-reduces (SUM) their gradients in ONE flat collective:
    applied to the head parameters so that every rank ends up with the FULL parameter gradient, like the
    reference's replicated head."""

    @staticmethod
    def forward(ctx, *ps):
        ctx.shapes = [p.shape for p in ps]
        return tuple(p.view_as(p) for p in ps)

    @staticmethod
    def backward(ctx, *gs):
        sizes = [int(torch.Size(s).numel()) for s in ctx.shapes]
        ref = next(g for g in gs if g is not None)
        extra = ops.EVENTS.pop("loss_out5_partial", None)          # partial losses of the captured step (see forward)
        n_extra = extra.numel() if extra is not None else 0
        if all(g is not None and g.dtype == ref.dtype for g in gs):
            # ONE gather kernel instead of a fill and a copy node per parameter (ten ~1.7 us nodes in a row at the end
            # of the captured step's critical path)
            parts = [g.reshape(-1) for g in gs]
            if extra is not None:
                parts.append(extra.detach().reshape(-1).to(ref.dtype))
            flat = torch.cat(parts)
        else:
            flat = torch.zeros(sum(sizes) + n_extra, dtype=ref.dtype, device=ref.device)
            if extra is not None:
                flat[sum(sizes):].copy_(extra.detach().reshape(-1))
            o = 0
            for g, n in zip(gs, sizes):
                if g is not None:
                    flat[o:o + n].copy_(g.reshape(-1))
                o += n
        ev = torch.cuda.Event()                 # everything that reads the step's operands is enqueued: the captured
        ev.record()                             # step starts its bank insert here, under the all-reduce
        ops.EVENTS["mlp_backward_done"] = ev
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if extra is not None:
            with torch.no_grad():
                extra.detach().copy_(flat[sum(sizes):].reshape(extra.shape))      # the losses every rank reports
        out, o = [], 0
        for shp, n in zip(ctx.shapes, sizes):
            out.append(flat[o:o + n].view(shp))
            o += n
        return tuple(out)


def sinkhorn_serial(B):
    """Global batch from which the forward contraction is ordered BEHIND the Sinkhorn chain instead of next to it
    (NR_SINKHORN_SERIAL = smallest such batch, 0 = never = default).  Measured on 8 B200 (global batch 1024,
    profiles/r2_n8_sinkhorn_serial_ab.txt): 624 steps/s serialised against 773 side by side — sharing the GPU costs
    both kernels less than the lost overlap."""
    v = int(os.environ.get("NR_SINKHORN_SERIAL", "0") or 0)
    return v > 0 and B >= v


_TEXT_STREAMS = {}


def _text_stream(device):
    st = _TEXT_STREAMS.get(device.index)
    if st is None:
        st = _TEXT_STREAMS[device.index] = torch.cuda.Stream(device=device)
    return st


class ShardedPrologue:
    """Everything of a sharded step that does not depend on the token-weight MLPs: the feature / mask / global-feature
    gathers, token preparation of the gathered batch and of the bank, centrality weights of the local rows, the
    replicated global similarity and its Sinkhorn duals.  The constructor allocates (and gathers the masks, which
    the preparation needs); the three run_* groups are independent and are enqueued on forked streams next to the
    MLP evaluations (modeling._sharded_losses) — the NCCL collectives overlap them."""

    def __init__(self, text_l, video_l, gt_l, gv_l, tm_l, vm_l, mb_feat_t, mb_feat_v, mb_mask_t, mb_mask_v, hp,
                 idx_l=None, bank_ring=None):
        cs, beta, k, tau, iters, wu, wn, wkl, prec, bprec = hp
        _req_cuda(text_l, video_l, gt_l, gv_l, mb_feat_t, mb_feat_v)
        self.hp = hp
        self.W, self.r = W, r = dist.get_world_size(), dist.get_rank()
        dev = text_l.device
        self.b = b = text_l.shape[0]
        self.B, self.lo = B, lo = W * b, r * b
        nt, nv, d = text_l.shape[1], video_l.shape[1], text_l.shape[2]
        f32 = dict(dtype=torch.float32, device=dev)
        self.text_l, self.video_l = _f32c(text_l.detach()), _f32c(video_l.detach())
        self.text = torch.empty(B, nt, d, **f32)
        self.video = torch.empty(B, nv, d, **f32)
        self.g_l = (_f32c(gt_l.detach()).reshape(b, d).contiguous(), _f32c(gv_l.detach()).reshape(b, d).contiguous())
        self.g2, self.v2 = torch.empty(B, d, **f32), torch.empty(B, d, **f32)
        x3 = prec == NR_PREC_BF16X3
        bf_ = prec in ops.TC_PRECISIONS or bprec in ops.TC_PRECISIONS
        fk_ = (prec in ops.TC_PRECISIONS and bprec == prec and ops.USE_FUSED_MAXSIM
               and ops.maxsim2_supported(nt, nv, d * (3 if x3 else 1)))
        if x3 and not fk_:
            raise RuntimeError("precision 'bf16x3' needs the fused two-direction kernel for these token counts")
        rx, ry = (ops.ROLE_X, ops.ROLE_Y) if x3 else (0, 0)
        if bf_ and not fk_ and ((b * nt) % 8 or (b * nv) % 8):
            # the one-direction bf16 backward reads row blocks of the transposed operand copies through 16-byte
            # aligned column offsets; decided from (b, Nt, Nv) alone so that EVERY rank raises before any collective
            raise RuntimeError(f"sharded head, bf16 one-direction kernels: per-rank batch x tokens must be a multiple "
                               f"of 8 (b={b}, Nt={nt}, Nv={nv}); use head_bwd_precision='fp32' or an even batch")
        # ---- exchange 0: ONE packed gather of everything small that exists before any kernel has run — the masks (the
        # operand preparation needs them), the global features (Sinkhorn starts from them) and the dataset indices of
        # the bank FIFO: [masks int64 b x (Nt+Nv) | globals f32 b x 2 x d | idx int64 b] as bytes
        idx_l = (idx_l if idx_l is not None else torch.zeros(b, dtype=torch.int64, device=dev)).reshape(b).to(torch.int64)
        pieces = [torch.cat([_mask(tm_l), _mask(vm_l)], dim=1).reshape(-1).view(torch.uint8),
                  torch.stack(self.g_l, 1).reshape(-1).view(torch.uint8), idx_l.contiguous().view(torch.uint8)]
        o1 = pieces[0].numel()
        o2 = o1 + pieces[1].numel()
        packed = _gather(torch.cat(pieces).unsqueeze(0))                                     # [W, bytes]
        masks = packed[:, :o1].view(torch.int64).reshape(B, nt + nv)
        self.tm, self.vm = masks[:, :nt].contiguous(), masks[:, nt:].contiguous()
        gall = packed[:, o1:o2].view(torch.float32).reshape(B, 2, d)
        self.g2.copy_(gall[:, 0]); self.v2.copy_(gall[:, 1])
        self.idx_all = packed[:, o2:].view(torch.int64).reshape(B)
        self.bank_static = bank_ring is not None
        if self.bank_static and (not fk_ or bank_ring.x3 != x3):
            raise RuntimeError("bank ring: needs the fused tensor-core path in the precision it was built for")
        self.mtm, self.mvm = (bank_ring.mask_t, bank_ring.mask_v) if self.bank_static else (_mask(mb_mask_t), _mask(mb_mask_v))
        self.bf, self.fusedk = bf, fk = bf_, fk_
        # Fused (bf16) path = "exchange" design: a rank contracts only ITS text rows against all videos (P = S[rows_r,
        # :]); the column block the video->text direction needs is assembled from the other ranks' P by an
        # all-to-all of [b,b] blocks.  Other ranks' text tokens are then needed by the bank FIFO only.
        self.a2a = fk
        if fk:
            self.T = None
            self.Tl = Prepared(self.text_l, bf16=True, colsum=True, mask=self.tm[lo:lo + b], defer=True, split=rx)
            self.tsum_l = torch.empty(1, d, **f32)
            self.tsum = torch.empty(W, d, **f32)
        else:
            self.T = Prepared(self.text, bf16=bf, colsum=True, mask=None, defer=True)
            self.Tl = self.T.block(lo, b)
        self.V = Prepared(self.video, bf16=bf, colsum=True, mask=self.vm if fk else None, defer=True, split=ry)
        if self.bank_static:
            self.MT, self.MV = bank_ring.MT, bank_ring.MV
        else:
            self.MT = Prepared(mb_feat_t, bf16=bf, mask=self.mtm if fk else None, defer=True, f32=not fk, split=rx)
            self.MV = Prepared(mb_feat_v, bf16=bf, mask=self.mvm if fk else None, defer=True, f32=not fk, split=ry)
        self.Vl = self.V.block(lo, b)
        self.t_rows = B * nt
        self.GG = torch.empty(2, B, B, **f32)              # [G ; G^T], replicated
        self.duals = torch.empty(4, B, **f32)
        self.lib_ws, self.nws = ops.sinkhorn_workspace(B, dev)
        self.mean = torch.empty(2, d, **f32); self.gn = torch.empty(2, b, d, **f32)
        self.ginv = torch.empty(2, b, **f32); self.w = torch.empty(2, b, **f32)
        self.global_done = None
        self.text_ready = None
        # set by callers that always run the backward (the captured step): the text gather then moves there
        self.defer_text_to_backward = False
        if bf:
            for P in (self.T if self.T is not None else self.Tl, self.V, self.MT, self.MV):
                P.alloc_transposed()

    def _centrality(self, partials, rows_total, i):
        b, d, cs = self.b, self.V.d, self.hp[0]
        _call("nr_centrality_fwd", _p(partials), partials.shape[0], rows_total, _p(self.g_l[i]), b, d, cs,
              _p(self.mean[i]), _p(self.gn[i]), _p(self.ginv[i]), _p(self.w[i]), _stream(), launches=2)

    def run_text_side(self):
        bprec = self.hp[9]
        if self.a2a:
            if not self.bank_static:
                self.MT.run()
            self.Tl.run()
            self.Tl.bwd_source(bprec)
            if not self.bank_static:
                self.MT.bwd_source(bprec)
            # column mean over ALL text tokens (modeling.py:419-424) from per-rank column sums: the [W, D] exchange
            # rides with the gather of the video token weights in the head forward (finish_text_centrality)
            torch.sum(self.Tl.partials, dim=0, keepdim=True, out=self.tsum_l)
            return                                                   # the text gather is deferred: gather_text_async()
        dist.all_gather_into_tensor(self.text, self.text_l)          # exchange 1a
        self.MT.run()
        self.T.run()
        if self.bf:
            self.MT.bwd_source(bprec); self.T.bwd_source(bprec)
        self._centrality(self.T.partials, self.T.rows, 0)

    def finish_text_centrality(self):
        self._centrality(self.tsum, self.t_rows, 0)

    def gather_text_async(self):
        """Exchange design only: other ranks' text tokens feed nothing but the memory-bank FIFO, so their gather is
        issued at the END of the forward on its own stream; consumers wait on `text_ready`."""
        main = torch.cuda.current_stream()
        side = _text_stream(main.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dist.all_gather_into_tensor(self.text, self.text_l)      # exchange 1a
            self.text_ready = torch.cuda.Event()
            self.text_ready.record(side)

    def run_video_side(self):
        bprec = self.hp[9]
        dist.all_gather_into_tensor(self.video, self.video_l)        # exchange 1b
        if not self.bank_static:
            self.MV.run()
        self.V.run()
        if self.bf:
            self.V.bwd_source(bprec)
            if not self.bank_static:
                self.MV.bwd_source(bprec)
        self._centrality(self.V.partials, self.V.rows, 1)

    def run_global(self):
        B, iters = self.B, int(self.hp[4])
        _call("nr_gram_f32", _p(self.g2), _p(self.v2), B, B, self.V.d, _p(self.GG[0]), _p(self.GG[1]), _stream())
        d_ = self.duals
        # in-step: 16 rows per CTA for the multi-CTA variants (the chain shares the GPU with the contraction); when the
        # contraction waits for the chain (sinkhorn_serial), the chain's own best setting
        rows = 0 if B < 256 else (8 if sinkhorn_serial(B) else 16)
        _call("nr_sinkhorn_ex", _p(self.GG[0]), _p(self.GG[1]), B, iters, _p(d_[0]), _p(d_[1]), _p(d_[2]), _p(d_[3]),
              _p(self.lib_ws), self.nws, rows, _stream())

    def run_forked(self):
        with ops.ForkJoin(2) as fj:
            with fj.on(1):
                self.run_global()            # first: its small gather must not queue behind the feature gathers
            with fj.on(0):
                self.run_video_side()
            self.run_text_side()
        return self


class ShardedHeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, text_l, video_l, gt_l, gv_l, tw_l, vw_l, tw_mb, vw_mb, logit_scale, tm_l, vm_l, mb_feat_t,
                mb_feat_v, mb_mask_t, mb_mask_v, hp, pro=None):
        cs, beta, k, tau, iters, wu, wn, wkl, prec, bprec = hp
        _req_cuda(text_l, video_l, gt_l, gv_l, tw_l, vw_l, tw_mb, vw_mb, logit_scale)
        ctx.set_materialize_grads(False)         # no zero-filled gradients for the gathered tensors / neighbour lists
        if pro is None:
            pro = ShardedPrologue(text_l, video_l, gt_l, gv_l, tm_l, vm_l, mb_feat_t, mb_feat_v, mb_mask_t,
                                  mb_mask_v, hp).run_forked()
        W, r, b, B, lo = pro.W, pro.r, pro.b, pro.B, pro.lo
        dev = text_l.device
        st = _stream()
        d = text_l.shape[-1]
        nt_, nv_ = text_l.shape[1], video_l.shape[1]
        text, video, tm, vm, mtm, mvm = pro.text, pro.video, pro.tm, pro.vm, pro.mtm, pro.mvm
        T, V, MT, MV, Tl, Vl = pro.T, pro.V, pro.MT, pro.MV, pro.Tl, pro.Vl
        fusedk = pro.fusedk
        # ---- exchange 2: token weights of the local rows (the only gather that needs the MLPs)
        a2a = pro.a2a
        if a2a:                                   # other ranks never touch this rank's text weights
            # [video token weights b x Nv | text column sums d] per rank in one gather
            pk = _gather(torch.cat([_f32c(vw_l).reshape(-1), pro.tsum_l.reshape(-1)]).unsqueeze(0))   # [W, b*Nv + d]
            vw = pk[:, :b * nv_].reshape(B, nv_)                                             # [B, Nv]
            pro.tsum.copy_(pk[:, b * nv_:])
            pro.finish_text_centrality()
            tw = _f32c(tw_l)                                                                 # [b, Nt]
            tw_lc = tw
        else:
            small = _gather(torch.cat([_f32c(tw_l), _f32c(vw_l)], dim=1))                    # [B, Nt+Nv]
            tw, vw = small[:, :nt_].contiguous(), small[:, nt_:].contiguous()
            tw_lc = tw[lo:lo + b]
        tw_mb, vw_mb = _f32c(tw_mb), _f32c(vw_mb)
        M = MT.r
        nt, nv = Tl.n, V.n
        vw_lc, tm_lc, vm_lc = vw[lo:lo + b], tm[lo:lo + b], vm[lo:lo + b]
        f32 = dict(dtype=torch.float32, device=dev)
        S_row = torch.empty(b, B, **f32); S_col = torch.empty(b, B, **f32)
        mbb = torch.empty(2, b, M, **f32)
        mb_t2v, mb_v2t = mbb[0], mbb[1]
        if fusedk:
            # ONE launch, 3 problems, every token pair multiplied once: S_row = S(text_l, video) [b,B] (and its
            # transpose [B,b], whose [b,b] row chunks are what the other ranks need) and the two bank blocks
            # the transposed block is written with a leading dimension of b + 2: chunk q (rows q*b..) is then the
            # [b, b+2] payload for rank q, whose two spare columns carry this rank's bank centralities (exchange 3)
            PT = torch.empty(B, b + 2, **f32)
            if sinkhorn_serial(B) and pro.global_done is not None:
                # large global batches: the Sinkhorn chain and the contraction exclude each other on an SM and both
                # crawl when they share the GPU; the chain starts ~0.1 ms earlier, so the contraction waits for it
                torch.cuda.current_stream().wait_event(pro.global_done)
                pro.global_done = None
            sv1, svA, svC = ops.maxsim2_fwd([
                dict(X=Tl, Y=V, wx=tw_lc, wy=vw, alpha=0.5, out=S_row, strides=(B, 1), out2=PT, strides2=(1, b + 2)),
                dict(X=Tl, Y=MV, wx=tw_lc, wy=vw_mb, alpha=0.5, out=mb_t2v, strides=(M, 1)),
                dict(X=MT, Y=Vl, wx=tw_mb, wy=vw_lc, alpha=0.5, out=mb_v2t, strides=(1, M))])
            p1, y1, p2, y2 = sv1          # (pmax_x, ystar, pmax_y, xstar) of each pair
            p3 = y3 = p4 = y4 = None
            pA, yA, pB, yB = svA
            pC, yC, pD, yD = svC
            c_l = torch.empty(2, b, **f32)
            _call("nr_row_mean", _p(mbb), M, 2 * b, M, _p(c_l), st)
            PT.view(W, b, b + 2)[:, :, b:] = c_l.t()
            # ---- exchange 2b + 3: S_col[v_l, (q, a)] = S[(q, a), lo + v_l] = chunk r of rank q's transposed block;
            #      the bank centrality of every sample (indexed by COLUMN in the neighbour loss) in the spare columns
            recv = _all_to_all_blocks(PT.view(W, b, b + 2))                                  # [q, v_l, a | c_q]
            S_col.view(b, W, b).copy_(recv[:, :, :b].permute(1, 0, 2))
            cb = recv[:, :, b:].permute(2, 0, 1).reshape(2, B).contiguous()                  # [c_t2v ; c_v2t]
        else:
            p1, y1 = _fwd_dir(prec, Tl, V, tw_lc, tm_lc, vm, S_row, B, 1, None, 0, 0, 0)      # H(text_l, video)
            p2, y2 = _fwd_dir(prec, V, Tl, vw, vm, tm_lc, S_row, 1, B, None, 0, 0, 1)         # H(video, text_l)^T
            p3, y3 = _fwd_dir(prec, Vl, T, vw_lc, vm_lc, tm, S_col, B, 1, None, 0, 0, 0)      # H(video_l, text)
            p4, y4 = _fwd_dir(prec, T, Vl, tw, tm, vm_lc, S_col, 1, B, None, 0, 0, 1)         # H(text, video_l)^T
            pA, yA = _fwd_dir(prec, Tl, MV, tw_lc, tm_lc, mvm, mb_t2v, M, 1, None, 0, 0, 0)
            pB, yB = _fwd_dir(prec, MV, Tl, vw_mb, mvm, tm_lc, mb_t2v, 1, M, None, 0, 0, 1)
            pD, yD = _fwd_dir(prec, Vl, MT, vw_lc, vm_lc, mtm, mb_v2t, M, 1, None, 0, 0, 0)
            pC, yC = _fwd_dir(prec, MT, Vl, tw_mb, mtm, vm_lc, mb_v2t, 1, M, None, 0, 0, 1)
        ctx.fusedk = fusedk
        if not fusedk:
            c_l = torch.empty(2, b, **f32)
            _call("nr_row_mean", _p(mbb), M, 2 * b, M, _p(c_l), st)
            # ---- exchange 3: bank centrality of every sample (indexed by COLUMN in the neighbour loss)
            cb = _gather(c_l.unsqueeze(0)).permute(1, 0, 2).reshape(2, B).contiguous()       # [c_t2v ; c_v2t]
        if pro.global_done is not None:                  # G, G^T and the Sinkhorn duals from the detached branch
            torch.cuda.current_stream().wait_event(pro.global_done)
            pro.global_done = None
        g2, v2, G, GT, duals = pro.g2, pro.v2, pro.GG[0], pro.GG[1], pro.duals
        mean, gn, ginv, w = pro.mean, pro.gn, pro.ginv, pro.w
        ls = _f32c(logit_scale).reshape(1)
        row_out = torch.zeros(2, 4, b, **f32)
        nbr = torch.empty(2, b, k, dtype=torch.int32, device=dev)
        saved = torch.empty(2, b, NR_NSAVE, **f32)
        sums = torch.empty(8, **f32)
        with ops.ForkJoin(1) as fj:
            _call("nr_row_losses_fwd", _p(S_row), B, _p(G[lo:lo + b]), B, _p(cb[1]), _p(w[0]), _p(duals[0][lo:lo + b]),
                  _p(duals[1]), b, B, lo, _p(ls), k, tau, tau, beta, ALL_LOSSES, _p(row_out[0]), _p(nbr[0]),
                  _p(saved[0]), _stream())
            with fj.on(0):
                _call("nr_row_losses_fwd", _p(S_col), B, _p(GT[lo:lo + b]), B, _p(cb[0]), _p(w[1]),
                      _p(duals[2][lo:lo + b]), _p(duals[3]), b, B, lo, _p(ls), k, tau, tau, beta, ALL_LOSSES,
                      _p(row_out[1]), _p(nbr[1]), _p(saved[1]), _stream())
        _call("nr_vec_sums", _p(row_out), 8, b, None, _p(sums), st)
        # ---- exchange 4: loss partial sums.  The backward never reads the reduced values (its upstream multipliers are
        #      constants), so the captured step does not put this collective between the forward and the backward
        #      exchange (measured at 8 ranks: 155 us of rank skew absorbed right there): the five partial losses ride
        #      on the all-reduce of the head-parameter gradients at the end of the backward (SumGradsAcrossRanks).
        piggy = bool(a2a and pro.defer_text_to_backward)
        if not piggy:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        m54 = _combine_matrix(B, wu, wn, wkl, dev)
        out5 = torch.empty(5, **f32)
        _call("nr_matvec_small", _p(m54), 5, 4, 0, _p(sums), _p(sums[4:]), _p(out5), st)
        if piggy:
            ops.EVENTS["loss_out5_partial"] = out5
        if a2a and not pro.defer_text_to_backward:
            pro.gather_text_async()
        ctx.pro = pro if (a2a and pro.defer_text_to_backward) else None
        ctx.defer_video_rs = bool(a2a and pro.defer_text_to_backward)
        ctx.hp, ctx.dims = hp, (W, r, b, B, lo, M, d, nt, nv)
        ctx.objs = (T, V, MT, MV, Tl, Vl)
        ctx.a2a = a2a
        ctx.save_for_backward(tw, vw, tw_mb, vw_mb, tm, vm, mtm, mvm, S_row, S_col, G, GT, cb, duals, w, ls, nbr, saved,
                              mean, gn, ginv, g2, v2, m54, p1, y1, p2, y2, p3, y3, p4, y4, pA, yA, pB, yB, pC, yC, pD, yD)
        ctx.gshape = (gt_l.shape, gv_l.shape)
        ctx.mark_non_differentiable(nbr, text, video, tm, vm)
        # the gathered batch is also what the memory-bank FIFO stores (reference modeling.py:309-310)
        return out5, nbr, text, video, tm, vm

    @staticmethod
    def backward(ctx, g5, _gn, _g1, _g2, _g3, _g4):
        (tw, vw, tw_mb, vw_mb, tm, vm, mtm, mvm, S_row, S_col, G, GT, cb, duals, w, ls, nbr, saved, mean, gn, ginv, g2,
         v2, m54, p1, y1, p2, y2, p3, y3, p4, y4, pA, yA, pB, yB, pC, yC, pD, yD) = ctx.saved_tensors
        cs, beta, k, tau, iters, wu, wn, wkl, prec, bprec = ctx.hp
        W, r, b, B, lo, M, d, nt, nv = ctx.dims
        T, V, MT, MV, Tl, Vl = ctx.objs
        dev = S_row.device
        st = _stream()
        f32 = dict(dtype=torch.float32, device=dev)
        gscale = torch.empty(4, **f32)
        _call("nr_matvec_small", _p(m54), 5, 4, 1, _p(_f32c(g5)), None, _p(gscale), st)
        # [dmean_text | dc_t2v | dc_v2t | dls | dw_t | dw_v]  (dmean first: it is read with 16-byte loads)
        z = torch.zeros(d + 2 * B + 1 + 2 * b, **f32)
        dcs, dls, dw = z[d:d + 2 * B + 1], z[d + 2 * B:d + 2 * B + 1], z[d + 2 * B + 1:].view(2, b)
        dc = z[d:d + 2 * B].view(2, B)
        dS_row = torch.empty(b, B, **f32); dS_col = torch.empty(b, B, **f32)
        dG1 = torch.empty(b, B, **f32); dG2 = torch.empty(b, B, **f32)
        _call("nr_row_losses_bwd", _p(S_row), B, _p(G[lo:lo + b]), B, _p(cb[1]), _p(w[0]), _p(duals[0][lo:lo + b]),
              _p(duals[1]), b, B, lo, _p(ls), k, tau, tau, beta, ALL_LOSSES, _p(nbr[0]), _p(saved[0]), _p(gscale),
              _p(dS_row), B, _p(dG1), B, _p(dc[1]), _p(dw[0]), _p(dls), st)
        _call("nr_row_losses_bwd", _p(S_col), B, _p(GT[lo:lo + b]), B, _p(cb[0]), _p(w[1]), _p(duals[2][lo:lo + b]),
              _p(duals[3]), b, B, lo, _p(ls), k, tau, tau, beta, ALL_LOSSES, _p(nbr[1]), _p(saved[1]), _p(gscale),
              _p(dS_col), B, _p(dG2), B, _p(dc[0]), _p(dw[1]), _p(dls), st)
        if ctx.a2a:
            dtext, dvideo, dgt, dgv, dtw_o, dvw_o, dtw_mb, dvw_mb = _backward_exchange(
                ctx, tw, vw, tw_mb, vw_mb, G, GT, w, mean, gn, ginv, g2, v2, p1, y1, p2, y2, pA, yA, pB, yB, pC, yC, pD,
                yD, z, dc, dw, dS_row, dS_col, dG1, dG2)
            ctx.objs = None
            gs_t, gs_v = ctx.gshape
            return (dtext, dvideo, dgt.reshape(gs_t), dgv.reshape(gs_v), dtw_o, dvw_o, dtw_mb, dvw_mb,
                    dls.reshape(()), None, None, None, None, None, None, None, None)
        # ---- exchange 4: gradients of the gathered bank-centrality vectors and of logit_scale (tiny)
        dist.all_reduce(dcs, op=dist.ReduceOp.SUM)
        dc_l = dc[:, lo:lo + b].contiguous()                           # this rank's samples
        # global similarity: dG has a row block (direction 1) and a column block (direction 2) on this rank
        dG = torch.zeros(B, B, **f32)
        dG[lo:lo + b] += dG1
        dG[:, lo:lo + b] += dG2.t()
        dg_all = dG @ v2                                               # partial over ranks
        dv_all = dG.t() @ g2
        dmean = torch.empty(2, d, **f32)
        dgl = torch.empty(2, b, d, **f32)
        _call("nr_centrality_bwd", _p(mean[0]), _p(gn[0]), _p(ginv[0]), _p(w[0]), _p(dw[0]), b, d, cs, T.rows,
              _p(dgl[0]), 0, _p(dmean[0]), st, launches=2)
        _call("nr_centrality_bwd", _p(mean[1]), _p(gn[1]), _p(ginv[1]), _p(w[1]), _p(dw[1]), b, d, cs, V.rows,
              _p(dgl[1]), 0, _p(dmean[1]), st, launches=2)
        dg_all[lo:lo + b] += dgl[0]
        dv_all[lo:lo + b] += dgl[1]
        # ---- token-pair products into full-size (gathered) gradient buffers
        dtn = torch.zeros(T.rows, d, **f32); dvn = torch.zeros(V.rows, d, **f32)
        dtw = torch.zeros(B, nt, **f32); dvw = torch.zeros(B, nv, **f32)
        dtw_mb = torch.zeros_like(tw_mb); dvw_mb = torch.zeros_like(vw_mb)
        dtn_l, dvn_l = dtn[lo * nt:(lo + b) * nt], dvn[lo * nv:(lo + b) * nv]
        dtw_l, dvw_l = dtw[lo:lo + b], dvw[lo:lo + b]
        tw_l, vw_l, tm_l, vm_l = tw[lo:lo + b], vw[lo:lo + b], tm[lo:lo + b], vm[lo:lo + b]
        vs, vld = V.bwd_source(bprec); ts, tld = T.bwd_source(bprec)
        vls, vlld = Vl.bwd_source(bprec); tls, tlld = Tl.bwd_source(bprec)
        mvs, mvld = MV.bwd_source(bprec); mts, mtld = MT.bwd_source(bprec)
        # (the fused two-direction kernels imply the exchange design, which returned above: only the one-direction
        # kernels are left here)
        if True:
            X, Y, Wg = "nr_maxsim_bwd_x", "nr_maxsim_bwd_y", "nr_maxsim_bwd_w"
            # H1 = H(text_l, video): dH1[a_l, bb] = .5 dS_row
            _call(X, bprec, _p(vs), vld, _p(tw_l), _p(tm_l), _p(vm), _p(y1), _p(dS_row), B, 1, 0.5, b, nt, B, nv, d, _p(dtn_l), st)
            _call(Y, bprec, _p(tls), tlld, _p(tw_l), _p(tm_l), _p(vm), _p(y1), _p(dS_row), B, 1, 0.5, b, nt, B, nv, d, _p(dvn), st)
            _call(Wg, _p(p1), _p(dS_row), B, 1, 0.5, b, nt, B, _p(dtw_l), st)
            # H2 = H(video, text_l): dH2[bb, a_l] = .5 dS_row[a_l, bb]
            _call(X, bprec, _p(tls), tlld, _p(vw), _p(vm), _p(tm_l), _p(y2), _p(dS_row), 1, B, 0.5, B, nv, b, nt, d, _p(dvn), st)
            _call(Y, bprec, _p(vs), vld, _p(vw), _p(vm), _p(tm_l), _p(y2), _p(dS_row), 1, B, 0.5, B, nv, b, nt, d, _p(dtn_l), st)
            _call(Wg, _p(p2), _p(dS_row), 1, B, 0.5, B, nv, b, _p(dvw), st)
            # H3 = H(video_l, text): dH3[v_l, a] = .5 dS_col
            _call(X, bprec, _p(ts), tld, _p(vw_l), _p(vm_l), _p(tm), _p(y3), _p(dS_col), B, 1, 0.5, b, nv, B, nt, d, _p(dvn_l), st)
            _call(Y, bprec, _p(vls), vlld, _p(vw_l), _p(vm_l), _p(tm), _p(y3), _p(dS_col), B, 1, 0.5, b, nv, B, nt, d, _p(dtn), st)
            _call(Wg, _p(p3), _p(dS_col), B, 1, 0.5, b, nv, B, _p(dvw_l), st)
            # H4 = H(text, video_l): dH4[a, v_l] = .5 dS_col[v_l, a]
            _call(X, bprec, _p(vls), vlld, _p(tw), _p(tm), _p(vm_l), _p(y4), _p(dS_col), 1, B, 0.5, B, nt, b, nv, d, _p(dtn), st)
            _call(Y, bprec, _p(ts), tld, _p(tw), _p(tm), _p(vm_l), _p(y4), _p(dS_col), 1, B, 0.5, B, nt, b, nv, d, _p(dvn_l), st)
            _call(Wg, _p(p4), _p(dS_col), 1, B, 0.5, B, nt, b, _p(dtw), st)
            # bank pairs of this rank's samples: dH = dc_l[a]/M broadcast over the bank rows (stride 0)
            sc = 0.5 / M
            _call(X, bprec, _p(mvs), mvld, _p(tw_l), _p(tm_l), _p(mvm), _p(yA), _p(dc_l[0]), 1, 0, sc, b, nt, M, nv, d, _p(dtn_l), st)
            _call(Wg, _p(pA), _p(dc_l[0]), 1, 0, sc, b, nt, M, _p(dtw_l), st)
            _call(Y, bprec, _p(mvs), mvld, _p(vw_mb), _p(mvm), _p(tm_l), _p(yB), _p(dc_l[0]), 0, 1, sc, M, nv, b, nt, d, _p(dtn_l), st)
            _call(Wg, _p(pB), _p(dc_l[0]), 0, 1, sc, M, nv, b, _p(dvw_mb), st)
            _call(X, bprec, _p(mts), mtld, _p(vw_l), _p(vm_l), _p(mtm), _p(yD), _p(dc_l[1]), 1, 0, sc, b, nv, M, nt, d, _p(dvn_l), st)
            _call(Wg, _p(pD), _p(dc_l[1]), 1, 0, sc, b, nv, M, _p(dvw_l), st)
            _call(Y, bprec, _p(mts), mtld, _p(tw_mb), _p(mtm), _p(vm_l), _p(yC), _p(dc_l[1]), 0, 1, sc, M, nt, b, nv, d, _p(dvn_l), st)
            _call(Wg, _p(pC), _p(dc_l[1]), 0, 1, sc, M, nt, b, _p(dtw_mb), st)
        dtext_all = torch.empty_like(T.xn); dvideo_all = torch.empty_like(V.xn)
        with ops.ForkJoin(1) as fj:
            T.backward(dtn, add_vec=dmean[0], out=dtext_all)
            with fj.on(0):
                V.backward(dvn, add_vec=dmean[1], out=dvideo_all)
        # ---- exchange 5: sum the partial gradients of the gathered tensors, keep this rank's rows
        dtext = _reduce_scatter(dtext_all, b)
        dvideo = _reduce_scatter(dvideo_all, b)
        small = _reduce_scatter(torch.cat([dg_all, dv_all, dtw, dvw], dim=1), b)            # one collective
        dgt, dgv = small[:, :d], small[:, d:2 * d]
        dtw_o, dvw_o = small[:, 2 * d:2 * d + nt], small[:, 2 * d + nt:]
        ctx.objs = None
        gs_t, gs_v = ctx.gshape
        return (dtext, dvideo, dgt.reshape(gs_t), dgv.reshape(gs_v), dtw_o, dvw_o, dtw_mb, dvw_mb, dls.reshape(()),
                None, None, None, None, None, None, None, None)


def _backward_exchange(ctx, tw, vw, tw_mb, vw_mb, G, GT, w, mean, gn, ginv, g2, v2, p1, y1, p2, y2, pA, yA, pB, yB, pC, yC,
                       pD, yD, z, dc, dw, dS_row, dS_col, dG1, dG2):
    """Backward of the exchange design (after the two row-loss backward launches): gradients w.r.t. S entries owned
    by other ranks travel back by the inverse all-to-all; text-side gradients are then complete locally, only the
    video side (all B videos on every rank) is reduce-scattered."""
    cs, beta, k, tau, iters, wu, wn, wkl, prec, bprec = ctx.hp
    W, r, b, B, lo, M, d, nt, nv = ctx.dims
    _T, V, MT, MV, Tl, Vl = ctx.objs
    dev = dS_row.device
    f32 = dict(dtype=torch.float32, device=dev)
    st = _stream()
    # centrality backward first: its text-side dmean rides on the same all_reduce as dc / d logit_scale
    dmean_v = torch.empty(d, **f32)
    dmean_t = z[:d]
    dgl = torch.empty(2, b, d, **f32)
    _call("nr_centrality_bwd", _p(mean[0]), _p(gn[0]), _p(ginv[0]), _p(w[0]), _p(dw[0]), b, d, cs, B * nt, _p(dgl[0]),
          0, _p(dmean_t), st, launches=2)
    _call("nr_centrality_bwd", _p(mean[1]), _p(gn[1]), _p(ginv[1]), _p(w[1]), _p(dw[1]), b, d, cs, V.rows, _p(dgl[1]),
          0, _p(dmean_v), st, launches=2)
    # ---- exchange 4 (ONE all-to-all): to rank q go the dS_col block whose text rows q owns, this rank's share of
    #      dc for q's samples, and this rank's d mean_text / d logit_scale (every rank's centrality weights read the
    #      global text mean); the receiver sums the tails over ranks:  dP[a, (r, v)] = dS_row[a, (r, v)] + recv[r, v, a]
    bb = b * b
    send = torch.empty(W, bb + 2 * b + d + 1, **f32)
    send[:, :bb].view(W, b, b).copy_(dS_col.view(b, W, b).permute(1, 0, 2))
    send[:, bb:bb + 2 * b].view(W, 2, b).copy_(dc.view(2, W, b).permute(1, 0, 2))
    send[:, bb + 2 * b:bb + 2 * b + d] = dmean_t
    send[:, bb + 2 * b + d:] = z[d + 2 * B:d + 2 * B + 1]
    recv = _all_to_all_blocks(send)                                                          # [r, (v, a) | tails]
    if ctx.pro is not None:
        # other ranks' text tokens feed only the bank FIFO at the very end of the step: their gather is queued behind
        # the exchange the backward waits for, and runs under the contraction
        ctx.pro.gather_text_async()
        ctx.pro = None
    tail = recv[:, bb:].sum(0)
    dc_l = tail[:2 * b].view(2, b).contiguous()                    # this rank's samples, summed over ranks
    dmean_t.copy_(tail[2 * b:2 * b + d])
    z[d + 2 * B:d + 2 * B + 1].copy_(tail[2 * b + d:])             # d logit_scale (returned by the caller)
    dP = dS_row.view(b, W, b) + recv[:, :bb].view(W, b, b).permute(2, 0, 1)
    dP = dP.view(b, B)
    # global similarity: dG has a row block (direction 1: dG1 [b,B] at rows lo..) and a column block (direction 2:
    # dG2^T) on this rank, so  dgT = dG gV  and  dgV = dG^T gT  are four [b,B] x [B,d] / [B,b] x [b,d] products
    # (4 b B d multiply-adds instead of 2 B^2 d on an assembled [B,B] matrix); partial over ranks.  They run on their
    # own branch, which stays open until the reduce-scatter that needs them.
    dg_all = torch.empty(B, d, **f32); dv_all = torch.empty(B, d, **f32)
    v2_l, g2_l = v2[lo:lo + b], g2[lo:lo + b]

    def global_path():
        mm = lambda A, tA, X, M_, K_, out, acc: _call("nr_matmul_f32", _p(A), B, tA, _p(X), d, M_, K_, d, _p(out), d, acc,
                                                      _stream())
        mm(dG2, 1, v2_l, B, b, dg_all, 0)                          # column block: dG[:, lo:lo+b] = dG2^T
        mm(dG1, 0, v2, b, B, dg_all[lo:lo + b], 1)                 # row block
        mm(dG1, 1, g2_l, B, b, dv_all, 0)
        mm(dG2, 0, g2, b, B, dv_all[lo:lo + b], 1)
        dg_all[lo:lo + b] += dgl[0]
        dv_all[lo:lo + b] += dgl[1]
    # ---- token-pair products: text rows complete locally, video rows partial over ranks.  One zero fill for every
    #      accumulation target: [dtn_l | dvn | dtw_l | dvw | dtw_mb | dvw_mb]  (token gradients first: 16-byte aligned)
    n_t, n_v = Tl.rows * d, V.rows * d
    sizes = [n_t, n_v, b * nt, B * nv, tw_mb.numel(), vw_mb.numel()]
    zz = torch.zeros(sum(sizes), **f32)
    offs = [0]
    for n_ in sizes:
        offs.append(offs[-1] + n_)
    dtn_l, dvn = zz[offs[0]:offs[1]].view(Tl.rows, d), zz[offs[1]:offs[2]].view(V.rows, d)
    dtw_l, dvw = zz[offs[2]:offs[3]].view(b, nt), zz[offs[3]:offs[4]].view(B, nv)
    dtw_mb, dvw_mb = zz[offs[4]:offs[5]].view_as(tw_mb), zz[offs[5]:offs[6]].view_as(vw_mb)
    dvn_l, dvw_l = dvn[lo * nv:(lo + b) * nv], dvw[lo:lo + b]
    vw_l = vw[lo:lo + b]
    dtext = torch.empty_like(Tl.xn); dvideo_all = torch.empty_like(V.xn)
    sc = 0.5 / M
    # The small dG products run on their own branch and are only joined by the LAST collective of the backward (the
    # global-feature gradients feed nothing inside the step): next to the contraction, whose persistent CTAs and the
    # ~1700 CTAs of the weight-gradient kernel hold every SM, they may start late without delaying anything.  The video
    # token-weight gradients, which gate the MLP backward, get their own small reduce-scatter right after the kernels.
    main = torch.cuda.current_stream()
    gstream = ops.ForkJoin(1, offset=6).side[0]
    gstream.wait_stream(main)
    with ops.ForkJoin(1) as fj:
        with fj.on(0):                           # weight gradients NEXT TO the contraction, not after it
            ops.maxsim2_bwd_w_multi([(p1, p2, dP, B, 1, 0.5, b, B, dtw_l, dvw),
                                     (pA, pB, dc_l[0], 1, 0, sc, b, M, dtw_l, dvw_mb),
                                     (pC, pD, dc_l[1], 0, 1, sc, M, b, dtw_mb, dvw_l)], nt, nv)
        ops.maxsim2_bwd_multi([
            (0, V, tw, vw, y1, y2, dP, B, 1, 0.5, b, B, dtn_l),
            (0, MV, tw, vw_mb, yA, yB, dc_l[0], 1, 0, sc, b, M, dtn_l),
            (1, Tl, tw, vw, y1, y2, dP, B, 1, 0.5, b, B, dvn),
            (1, MT, tw_mb, vw_l, yC, yD, dc_l[1], 0, 1, sc, M, b, dvn_l)], nt, nv, d)
        with torch.cuda.stream(gstream):         # enqueued after the contraction: a small kernel that grabs an SM
            global_path()                        # first would delay one of its statically scheduled CTAs
        # this branch outlives the function in the captured step: everything it reads or writes must stay allocated
        # until it has run (the caching allocator otherwise hands the blocks to the next main-stream allocation)
        for t_ in (dG1, dG2, v2, g2, dgl, dg_all, dv_all):
            t_.record_stream(gstream)
    dvw_o = _reduce_scatter(dvw, b)                                                          # [b, Nv]
    with ops.ForkJoin(1) as fj:
        Tl.backward(dtn_l, add_vec=dmean_t, out=dtext)
        with fj.on(0):
            V.backward(dvn, add_vec=dmean_v, out=dvideo_all)
    # ---- exchange 5: sum the partial video-side feature gradients and the global-feature gradients, keep this rank's
    #      rows.  Nothing of the step reads them; in the captured step (which joins `video_grad_ready` at its end) the
    #      two collectives therefore run from a side stream, under the MLP backward, instead of holding the main stream
    if ctx.defer_video_rs:
        side = _text_stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        side.wait_stream(gstream)
        with torch.cuda.stream(side):
            dvideo = _reduce_scatter(dvideo_all, b)
            small = _reduce_scatter(torch.cat([dg_all, dv_all], dim=1), b)
            ev = torch.cuda.Event()
            ev.record(side)
        for t_ in (dvideo_all, dg_all, dv_all):
            t_.record_stream(side)
        ops.EVENTS["video_grad_ready"] = ev
    else:
        main.wait_stream(gstream)
        dvideo = _reduce_scatter(dvideo_all, b)
        small = _reduce_scatter(torch.cat([dg_all, dv_all], dim=1), b)
    return dtext, dvideo, small[:, :d], small[:, d:2 * d], dtw_l, dvw_o, dtw_mb, dvw_mb


def sharded_head(text_l, video_l, gt_l, gv_l, tw_l, vw_l, tw_mb, vw_mb, logit_scale, tm_l, vm_l, mb_feat_t, mb_feat_v,
                 mb_mask_t, mb_mask_v, *, centrality_scale, beta, num_neighbors, temperature, uniform_weight,
                 neighbor_weight, kl_weight, precision="bf16", bwd_precision=None, iters=50, prologue=None):
    hp = head_hparams(centrality_scale, beta, num_neighbors, temperature, uniform_weight, neighbor_weight, kl_weight,
                      precision, bwd_precision, iters)
    return ShardedHeadFunction.apply(text_l, video_l, gt_l, gv_l, tw_l, vw_l, tw_mb, vw_mb, logit_scale, tm_l, vm_l,
                                     mb_feat_t, mb_feat_v, mb_mask_t, mb_mask_v, hp, prologue)
