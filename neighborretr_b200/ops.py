"""torch.autograd.Functions of the retrieval head on top of the C ABI (include/nrhead.h).

PyTorch is plumbing here: device buffers, the current CUDA stream, autograd graph edges.  All arithmetic
of the path runs in libnrhead.so kernels.  Every op refuses CPU tensors — there is no fallback.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from ._lib import (NR_LOSS_CENTRALITY, NR_LOSS_KL, NR_LOSS_NEIGHBOR, NR_LOSS_UNIFORM, NR_NSAVE, NR_PREC_BF16,
                   NR_PREC_BF16X3, NR_PREC_FP32, check)

# "bf16x3": fp32-accurate tensor-core mode — every normalised token value is split into hi + lo bf16 parts and the
# contraction runs hi.hi + lo.hi + hi.lo on the same tcgen05 kernels with K = 3d (forward) / three routing jobs per
# pair (backward); only the fused two-direction kernels implement it
PRECISIONS = {"fp32": NR_PREC_FP32, "bf16": NR_PREC_BF16, "bf16x3": NR_PREC_BF16X3}
TC_PRECISIONS = (NR_PREC_BF16, NR_PREC_BF16X3)
ROLE_X, ROLE_Y = 1, 2
# kernel launches issued through the C ABI since the last reset (bench.py reports it as gpu_launches)
LAUNCHES = {"count": 0}
# named CUDA events a later stage may wait on instead of the whole stream ("mlp_backward_done": sharded.py)
EVENTS = {}


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("neighborretr_b200: CUDA tensors required (no CPU fallback exists)")


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _mask(t):
    if t is None:
        return None
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t.contiguous()


_SIDE_STREAMS = {}
_SIDE_STREAMS_HP = {}


class ForkJoin:
    """Independent launch groups on side streams, joined back into the current stream on exit:

        with ForkJoin(2) as fj:
            ...                       # current stream
            with fj.on(0): ...        # side stream 0 (ordered after everything enqueued before the fork)
            with fj.on(1): ...

    Under CUDA-graph capture the side streams become parallel branches of the graph.  Callers allocate every buffer
    a branch touches BEFORE the fork, on the main stream (the caching allocator tracks one stream per block), and — for
    a DETACHED branch — keep a reference to every such buffer until the branch has been joined: a block freed when its
    Python owner goes out of scope is handed to the next main-stream allocation, which the still-running branch then
    races with (fused.HeadFunction keeps `row_out` / `sums` in ctx for exactly this reason)."""

    def __init__(self, n, offset=0, high=()):
        """offset: use side streams [offset, offset + n) — a fork nested inside another fork's main section must not
        reuse the outer fork's streams.  high: branch indices that run on HIGH-PRIORITY streams: small kernels that
        would otherwise queue behind the pending CTAs of a persistent kernel launched before them."""
        self.main = torch.cuda.current_stream()
        lst = _SIDE_STREAMS.get(self.main.device.index, [])
        while len(lst) < offset + n:
            lst.append(torch.cuda.Stream(device=self.main.device))
        _SIDE_STREAMS[self.main.device.index] = lst
        self.side = list(lst[offset:offset + n])
        for i in high:
            key = (self.main.device.index, offset + i)
            st = _SIDE_STREAMS_HP.get(key)
            if st is None:
                st = _SIDE_STREAMS_HP[key] = torch.cuda.Stream(device=self.main.device, priority=-1)
            self.side[i] = st
        self.detached = set()

    def __enter__(self):
        for s in self.side:
            s.wait_stream(self.main)
        return self

    def on(self, i):
        return torch.cuda.stream(self.side[i])

    def detach(self, i):
        """Do not join side stream i at exit: returns an event marking the end of what it has been given so far;
        the caller makes the consuming stream wait on it later (the branch keeps overlapping what follows)."""
        ev = torch.cuda.Event()
        ev.record(self.side[i])
        self.detached.add(i)
        return ev

    def __exit__(self, *exc):
        for i, s in enumerate(self.side):
            if i not in self.detached:
                self.main.wait_stream(s)
        return False


_HP_STREAMS = {}


def hp_stream(device=None):
    """One high-priority side stream per device for SMALL kernels that gate a later stage while a persistent kernel
    holds every SM: pending CTAs of a higher-priority stream are placed first whenever an SM has room, instead of
    queueing behind the thousands of CTAs of a concurrent low-urgency kernel."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    st = _HP_STREAMS.get(dev.index)
    if st is None:
        st = _HP_STREAMS[dev.index] = torch.cuda.Stream(device=dev, priority=-1)
    return st


def prezeroed(numel, device, offset):
    """fp32 zeros for accumulators of a LATER stage, filled on a side stream now: inside a captured step the fill is a
    parallel branch instead of a node on the critical path right before its consumer.  Only while capturing (returns
    None otherwise: the consumer then allocates its zeros itself).  Returns (flat tensor, event to wait on)."""
    if not torch.cuda.is_current_stream_capturing():
        return None
    z = torch.empty(int(numel), dtype=torch.float32, device=device)        # allocated on the main stream
    with ForkJoin(1, offset=offset) as fj:
        with fj.on(0):
            z.zero_()
        ev = fj.detach(0)
    return z, ev


class _KernelTimer:
    """CUDA-event timing of one C-ABI entry point on the launching stream (bench.py roofline).

    An event pair around a launch measures the kernel only if the GPU is still busy when the first event is
    enqueued; otherwise the host time between the two record() calls (building the argument block, encoding tensor
    maps, the launch itself: 10-20 us through ctypes) is counted as kernel time.  `hold(stream)` therefore parks the
    stream on a spin kernel first, so that the host runs ahead of the device through the whole timed step."""

    def __init__(self):
        self.name, self.events = None, []

    @staticmethod
    def hold(ms=3.0):
        torch.cuda._sleep(int(ms * 1.9e6))          # cycles at ~1.9 GHz

    def enable(self, name):
        self.name, self.events = name, []

    def disable(self):
        self.name, self.events = None, []

    def collect(self):
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for _, a, b in self.events)
        return {"ms": ms, "launches": len(self.events)}

    def table(self):
        """Per-entry-point totals (name -> (calls, ms)) when enabled with "*"."""
        torch.cuda.synchronize()
        out = {}
        for n, a, b in self.events:
            c, t = out.get(n, (0, 0.0))
            out[n] = (c + 1, t + a.elapsed_time(b))
        return out


KERNEL_TIMER = _KernelTimer()


def _call(name, *args, launches=1):
    fn = getattr(_lib.load(), name)
    if KERNEL_TIMER.name == name or KERNEL_TIMER.name == "*":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        KERNEL_TIMER.events.append((name, e0, e1))
    else:
        rc = fn(*args)
    check(rc, name)
    LAUNCHES["count"] += launches


# ------------------------------------------------------------------------------------------------
# token preparation
# ------------------------------------------------------------------------------------------------
class Prepared:
    """L2-normalised tokens of one modality: fp32 copy (backward / fp32 mode), optional bf16 operand
    copy (tensor-core mode), inverse norms and per-CTA column-sum partials."""

    __slots__ = ("xn", "xn_bf16", "xnT_bf16", "inv_norm", "partials", "rows", "n", "d", "r", "_parent", "_lo", "mask",
                 "_x", "_t_ready", "device", "split")

    def __init__(self, x, bf16=False, colsum=False, normalize=True, mask=None, defer=False, f32=True, split=0):
        """mask [r, n] int64 (optional): masked tokens become zero rows of the bf16 operand copy (and of its
        transposed copy) and take no max-sim gradient in backward(); required by the fused two-direction kernel.
        split: 0, or ROLE_X / ROLE_Y — the bf16 operand copy is the split-bf16 operand [r, n, 3d] of the bf16x3
        mode for that side of the fused kernel (nr_prep_tokens_split)."""
        _req_cuda(x)
        x = _f32c(x)
        self.r, self.n, self.d = x.shape
        self.rows = self.r * self.n
        self.xnT_bf16 = None
        self._parent, self._lo = None, 0
        self._x, self._t_ready = None, True
        self.mask = _mask(mask)
        self.split = int(split)
        dev = x.device
        self.device = dev
        if not normalize:        # global_level: raw dot products (reference modeling.py:525), fp32 only
            self.xn, self.xn_bf16, self.inv_norm, self.partials = x, None, None, None
            return
        # f32=False: no fp32 normalised copy (operands that never take a gradient in the bf16 mode, i.e. the memory
        # bank: saves 4 of the 6 bytes the preparation writes per element)
        self.xn = torch.empty_like(x) if f32 else None
        self.device = dev
        oshape = (self.r, self.n, 3 * self.d) if self.split else x.shape
        self.xn_bf16 = torch.empty(oshape, dtype=torch.bfloat16, device=dev) if (bf16 or self.split) else None
        self.inv_norm = torch.empty(self.rows, dtype=torch.float32, device=dev)
        npart = _lib.load().nr_prep_partials(self.rows)
        self.partials = torch.empty(npart, self.d, dtype=torch.float32, device=dev) if colsum else None
        self._x = x
        if not defer:
            self.run()

    def run(self):
        """Enqueue the preparation kernel on the current stream (buffers were allocated by the constructor)."""
        if self.split:
            _call("nr_prep_tokens_split", _p(self._x), self.rows, self.d, _p(self.xn), _p(self.xn_bf16), self.split,
                  _p(self.inv_norm), _p(self.partials), _p(self.mask), _stream())
        else:
            _call("nr_prep_tokens", _p(self._x), self.rows, self.d, _p(self.xn), _p(self.xn_bf16), _p(self.inv_norm),
                  _p(self.partials), _p(self.mask), _stream())
        self._x = None
        return self

    @property
    def kd(self):
        """Contraction length of the bf16 operand copy (3d for the split operand)."""
        return 3 * self.d if self.split else self.d

    def alloc_transposed(self):
        """Allocate (not fill) the transposed bf16 copy on the current stream; bwd_source() fills it on first use."""
        if self.xnT_bf16 is None and self.xn_bf16 is not None and self._parent is None:
            ld = (self.rows + 7) // 8 * 8
            self.xnT_bf16 = torch.empty(self.kd, ld, dtype=torch.bfloat16, device=self.xn_bf16.device)
            self._t_ready = False

    def operand(self, prec):
        return self.xn_bf16 if prec in TC_PRECISIONS else self.xn

    def bwd_source(self, prec):
        """(pointer tensor, ld) of this modality as the SOURCE operand of a backward contraction: fp32 tokens,
        or the transposed bf16 copy [d, ld] built on first use."""
        if prec not in TC_PRECISIONS:
            return self.xn, 0
        parent = getattr(self, "_parent", None)
        if parent is not None:          # block view: columns [lo*n, (lo+rows)) of the parent's transposed copy
            pt, ld = parent.bwd_source(prec)
            if (self._lo * self.n) % 8:
                raise RuntimeError("row block offset must keep the transposed operand 16-byte aligned")
            return pt[:, self._lo * self.n:], ld
        if self.xnT_bf16 is None:
            self.alloc_transposed()
        if not self._t_ready:
            _call("nr_transpose_tokens_bf16", _p(self.xn_bf16), self.rows, self.kd, _p(self.xnT_bf16),
                  self.xnT_bf16.shape[1], _stream())
            self._t_ready = True
        return self.xnT_bf16, self.xnT_bf16.shape[1]

    def block(self, lo, n):
        """View of samples [lo, lo+n) (a rank's rows of the gathered batch) usable wherever a Prepared is:
        same storage, pointer offsets only."""
        v = Prepared.__new__(Prepared)
        v.r, v.n, v.d, v.rows = n, self.n, self.d, n * self.n
        v.xn = self.xn[lo:lo + n] if self.xn is not None else None
        v.device = self.device
        v.xn_bf16 = self.xn_bf16[lo:lo + n] if self.xn_bf16 is not None else None
        v.inv_norm = self.inv_norm[lo * self.n:(lo + n) * self.n] if self.inv_norm is not None else None
        v.partials = None
        v.xnT_bf16 = None
        v._x, v._t_ready = None, True
        v._parent, v._lo = self, lo
        v.split = self.split
        v.mask = self.mask[lo:lo + n] if self.mask is not None else None
        return v

    def hi_lo_sources(self):
        """Split operand only: (hi^T, lo^T, ld) views of the transposed copy [3d, ld] (segments: role X = hi|lo|hi,
        role Y = hi|hi|lo)."""
        t, ld = self.bwd_source(NR_PREC_BF16X3)
        d = self.d
        lo = t[d:2 * d] if self.split == ROLE_X else t[2 * d:3 * d]
        return t[:d], lo, ld

    def backward(self, dxn, add_vec=None, out=None):
        if self.inv_norm is None:
            return dxn
        dx = torch.empty_like(self.xn) if out is None else out
        _call("nr_prep_tokens_bwd", _p(self.xn), _p(self.inv_norm), _p(dxn), _p(add_vec), _p(self.mask), self.rows,
              self.d, _p(dx), 0, _stream())
        return dx


# ------------------------------------------------------------------------------------------------
# max-sim (one direction) raw helpers
# ------------------------------------------------------------------------------------------------
def _maxsim_dir_fwd(prec, X: Prepared, Y: Prepared, wx, mx, my, alpha, out, sr, sc, out2, sr2, sc2, accumulate,
                    keep):
    dev = X.device
    pmax = torch.empty(X.r, Y.r, X.n, dtype=torch.float32, device=dev) if keep else None
    ystar = torch.empty(X.r, Y.r, X.n, dtype=torch.uint8, device=dev) if keep else None
    _call("nr_maxsim_fwd", prec, _p(X.operand(prec)), _p(Y.operand(prec)), _p(wx), _p(mx), _p(my), X.r, X.n, Y.r,
          Y.n, X.d, alpha, _p(out), sr, sc, _p(out2), sr2, sc2, accumulate, _p(pmax), _p(ystar), _stream())
    return pmax, ystar


def maxsim2_supported(nx, ny, d):
    return bool(_lib.load().nr_maxsim2_supported(nx, ny, d))


_TILE_WS = {}


def _tile_workspace(dev):
    """Persistent, self-resetting scheduler counters: one zero-initialised 16-byte buffer per (device, stream)."""
    cap = torch.cuda.is_current_stream_capturing()
    new = lambda: torch.zeros(4 + 2 * 16 * 256, dtype=torch.int32, device=dev)        # counters + debug timeline
    key = (dev.index, "capture" if cap else torch.cuda.current_stream(dev).cuda_stream)
    ws = _TILE_WS.get(key)
    if ws is None:
        ws = _TILE_WS[key] = new()
    if not cap and (dev.index, "capture") not in _TILE_WS:
        # captured steps (one launch per step, replays never overlap on a device) share a buffer created HERE, outside
        # any capture: allocated while capturing, its zero fill would become a node in front of every replayed launch
        _TILE_WS[(dev.index, "capture")] = new()
    return ws


def maxsim2_fwd(problems, keep=True):
    """Fused two-direction max-sim (nr_maxsim2_fwd) of up to 4 problems in ONE launch.  Each problem is a dict with
    X, Y (Prepared with zeroed masked tokens), wx, wy, alpha, out, strides=(sr, sc) and optionally out2, strides2.
    Returns per problem (pmax_x, ystar, pmax_y, xstar) (None if not keep)."""
    arr = (_lib.MaxSim2Problem * len(problems))()
    saved = []
    nx, ny, d = problems[0]["X"].n, problems[0]["Y"].n, problems[0]["X"].kd
    for i, q in enumerate(problems):
        X, Y = q["X"], q["Y"]
        if (X.n, Y.n, X.kd, Y.kd) != (nx, ny, d, d):
            raise RuntimeError("maxsim2_fwd: all problems of a launch must share (Nx, Ny, d)")
        if X.xn_bf16 is None or Y.xn_bf16 is None:
            raise RuntimeError("maxsim2_fwd: bf16 operand copies required (Prepared(..., bf16=True, mask=...))")
        if (X.split, Y.split) not in ((0, 0), (ROLE_X, ROLE_Y)):
            raise RuntimeError("maxsim2_fwd: split operands must be prepared as (ROLE_X, ROLE_Y)")
        dev = X.device
        if keep:
            sv = (torch.empty(X.r, Y.r, nx, dtype=torch.float32, device=dev),
                  torch.empty(X.r, Y.r, nx, dtype=torch.uint8, device=dev),
                  torch.empty(X.r, Y.r, ny, dtype=torch.float32, device=dev),
                  torch.empty(X.r, Y.r, ny, dtype=torch.uint8, device=dev))
        else:
            sv = (None, None, None, None)
        saved.append(sv)
        a = arr[i]
        a.x_bf16, a.y_bf16 = X.xn_bf16.data_ptr(), Y.xn_bf16.data_ptr()
        a.wx, a.wy = q["wx"].data_ptr(), q["wy"].data_ptr()
        a.Rx, a.Ry, a.alpha = X.r, Y.r, float(q["alpha"])
        a.out, (a.out_sr, a.out_sc) = q["out"].data_ptr(), q["strides"]
        o2 = q.get("out2")
        a.out2 = o2.data_ptr() if o2 is not None else None
        a.out2_sr, a.out2_sc = q.get("strides2", (0, 0))
        a.pmax_x, a.ystar, a.pmax_y, a.xstar = [t.data_ptr() if t is not None else None for t in sv]
    ws = _tile_workspace(dev)                                   # counters of the dynamic tile scheduler
    if problems[0]["X"].split:      # split-bf16 operands: exact-order column keys
        _call("nr_maxsim2_fwd_ex", ctypes.cast(arr, ctypes.c_void_p), len(problems), nx, ny, d, _p(ws), 1, _stream())
    else:
        _call("nr_maxsim2_fwd", ctypes.cast(arr, ctypes.c_void_p), len(problems), nx, ny, d, _p(ws), _stream())
    return saved


def maxsim2_bwd_multi(jobs, nx, ny, d):
    """All token-gradient contractions of a step in ONE launch (nr_maxsim2_bwd).  jobs: tuples
    (side, src Prepared, wx, wy, ystar, xstar, g, g_sr, g_sc, scale, rx, ry, dst); side 0: dst = X-token gradient
    (src = the Y tokens), side 1: dst = Y-token gradient (src = the X tokens).  Jobs with the same dst are
    accumulated in one pass."""
    flat = []
    for job in jobs:
        src = job[1]
        if src.split:       # bf16x3: (hi tile, hi^T), (lo tile, hi^T), (hi tile, lo^T) on the same destination
            hiT, loT, ld = src.hi_lo_sources()
            flat += [(job, hiT, ld, 2), (job, hiT, ld, 3), (job, loT, ld, 2)]
        else:
            s, ld = src.bwd_source(NR_PREC_BF16)
            flat.append((job, s, ld, 0))
    arr = (_lib.MaxSim2BwdJob * len(flat))()
    for i, ((side, src, wx, wy, ystar, xstar, g, g_sr, g_sc, scale, rx, ry, dst), s, ld, part) in enumerate(flat):
        a = arr[i]
        a.side, a.srcT, a.src_ld = side, s.data_ptr(), ld
        a.wx, a.wy, a.ystar, a.xstar = wx.data_ptr(), wy.data_ptr(), ystar.data_ptr(), xstar.data_ptr()
        a.dH, a.dh_sr, a.dh_sc, a.dh_scale = g.data_ptr(), g_sr, g_sc, float(scale)
        a.Rx, a.Ry, a.dst, a.part = rx, ry, dst.data_ptr(), part
    _call("nr_maxsim2_bwd", ctypes.cast(arr, ctypes.c_void_p), len(flat), nx, ny, d, _stream())


def maxsim2_bwd(side, src, wx, wy, ystar, xstar, g, g_sr, g_sc, scale, rx, nx, ry, ny, d, dst):
    """dst += (routing matrix of the fused max-sim, or its transpose) x the tokens of `src` (a Prepared whose
    transposed bf16 copy is the staged operand); side 0: dst = X-token gradient, side 1: Y-token gradient."""
    maxsim2_bwd_multi([(side, src, wx, wy, ystar, xstar, g, g_sr, g_sc, scale, rx, ry, dst)], nx, ny, d)


def maxsim2_bwd_w(pmax_x, pmax_y, g, g_sr, g_sc, scale, rx, nx, ry, ny, dwx, dwy):
    _call("nr_maxsim2_bwd_w", _p(pmax_x), _p(pmax_y), _p(g), g_sr, g_sc, scale, rx, nx, ry, ny, _p(dwx), _p(dwy),
          _stream())


def maxsim2_bwd_w_multi(jobs, nx, ny):
    """jobs: (pmax_x, pmax_y, g, g_sr, g_sc, scale, rx, ry, dwx, dwy) tuples (dwx / dwy nullable; jobs without any
    requested gradient are dropped) -> ONE launch."""
    jobs = [j for j in jobs if j[8] is not None or j[9] is not None]
    if not jobs:
        return
    arr = (_lib.MaxSim2BwdWJob * len(jobs))()
    for i, (px, py, g, g_sr, g_sc, scale, rx, ry, dwx, dwy) in enumerate(jobs):
        a = arr[i]
        a.pmax_x, a.pmax_y, a.dH = px.data_ptr(), py.data_ptr(), g.data_ptr()
        a.dh_sr, a.dh_sc, a.dh_scale, a.Rx, a.Ry = g_sr, g_sc, float(scale), rx, ry
        a.dwx = dwx.data_ptr() if dwx is not None else None
        a.dwy = dwy.data_ptr() if dwy is not None else None
    _call("nr_maxsim2_bwd_w_multi", ctypes.cast(arr, ctypes.c_void_p), len(jobs), nx, ny, _stream())


# the fused two-direction kernels are the bf16 path; tests flip this to compare against the one-direction kernels
USE_FUSED_MAXSIM = True


def _fused_orientation(nt, nv, d):
    """Which modality plays X in the fused kernel (its column-direction butterfly wants Nx % 4 == 0):
    False = text, True = video, None = neither orientation is supported."""
    if not USE_FUSED_MAXSIM:
        return None
    if maxsim2_supported(nt, nv, d):
        return False
    if maxsim2_supported(nv, nt, d):
        return True
    return None


class MaxSimFunction(torch.autograd.Function):
    """local_level's token-pair part (reference modeling.py:495-512) given the token weights:
    returns (S, S^T) as two contiguous tensors.  bwd_prec selects the arithmetic of the backward
    contraction independently of the forward one."""

    @staticmethod
    def forward(ctx, text_feat, video_feat, tw, vw, text_mask, video_mask, prec, bwd_prec, normalize=True):
        _req_cuda(text_feat, video_feat, tw, vw, text_mask, video_mask)
        tw, vw = _f32c(tw), _f32c(vw)
        tm, vm = _mask(text_mask), _mask(video_mask)
        if not normalize:
            prec = bwd_prec = NR_PREC_FP32
        x3 = prec == NR_PREC_BF16X3
        if x3:
            bwd_prec = NR_PREC_BF16X3
        need_bf16 = prec == NR_PREC_BF16 or bwd_prec == NR_PREC_BF16
        keep = any(ctx.needs_input_grad[:4])
        swap = None
        if prec in TC_PRECISIONS and bwd_prec == prec:
            swap = _fused_orientation(text_feat.shape[1], video_feat.shape[1], text_feat.shape[2] * (3 if x3 else 1))
        if x3 and swap is None:
            raise RuntimeError("precision 'bf16x3' needs the fused two-direction kernel (token counts 4..128 with one "
                               "side a multiple of 4 and the other in {4,8,12,16,24,32,48,64}, d % 64 == 0)")
        ctx.fused = swap is not None
        if ctx.fused:
            # one launch: both directions from the same accumulator tile, masks folded into the operand copies
            T = Prepared(text_feat, bf16=True, mask=tm, split=(ROLE_Y if swap else ROLE_X) if x3 else 0)
            V = Prepared(video_feat, bf16=True, mask=vm, split=(ROLE_X if swap else ROLE_Y) if x3 else 0)
            A, B = T.r, V.r
            S = torch.empty(A, B, dtype=torch.float32, device=T.xn.device)
            ST = torch.empty(B, A, dtype=torch.float32, device=T.xn.device)
            if swap:
                q = dict(X=V, Y=T, wx=vw, wy=tw, alpha=0.5, out=S, strides=(1, B), out2=ST, strides2=(A, 1))
            else:
                q = dict(X=T, Y=V, wx=tw, wy=vw, alpha=0.5, out=S, strides=(B, 1), out2=ST, strides2=(1, A))
            sv = maxsim2_fwd([q], keep=keep)[0]
            ctx.T, ctx.V, ctx.prec, ctx.swap = T, V, bwd_prec, swap
            ctx.save_for_backward(tw, vw, tm, vm, *sv)
            return S, ST
        T = Prepared(text_feat, bf16=need_bf16, normalize=normalize)
        V = Prepared(video_feat, bf16=need_bf16, normalize=normalize)
        A, B = T.r, V.r
        dev = T.xn.device
        S = torch.empty(A, B, dtype=torch.float32, device=dev)
        ST = torch.empty(B, A, dtype=torch.float32, device=dev)
        p1, y1 = _maxsim_dir_fwd(prec, T, V, tw, tm, vm, 0.5, S, B, 1, ST, 1, A, 0, keep)
        p2, y2 = _maxsim_dir_fwd(prec, V, T, vw, vm, tm, 0.5, S, 1, B, ST, A, 1, 1, keep)
        ctx.T, ctx.V, ctx.prec = T, V, bwd_prec
        ctx.save_for_backward(tw, vw, tm, vm, p1, y1, p2, y2)
        return S, ST

    @staticmethod
    def backward(ctx, dS, dST):
        tw, vw, tm, vm, p1, y1, p2, y2 = ctx.saved_tensors
        T, V, prec = ctx.T, ctx.V, ctx.prec
        A, B = T.r, V.r
        dev = T.xn.device
        st = _stream()
        # total gradient of S: dS + dST^T
        if dS is None and dST is None:
            return (None,) * 9
        if dST is None:
            g = _f32c(dS)
        else:
            g = torch.empty(A, B, dtype=torch.float32, device=dev)
            a = _f32c(dS) if dS is not None else None
            _call("nr_transpose_add", _p(a), B, _p(_f32c(dST)), A, _p(g), B, A, B, 1.0, 1.0, st)
        need_t, need_v, need_tw, need_vw = ctx.needs_input_grad[:4]
        dtn = torch.zeros_like(T.xn) if need_t else None
        dvn = torch.zeros_like(V.xn) if need_v else None
        dtw = torch.zeros_like(tw) if need_tw else None
        dvw = torch.zeros_like(vw) if need_vw else None
        if ctx.fused:
            if ctx.swap:      # X = video, Y = text: g(rx=b, ry=a) = g[a,b]
                dims = (B, V.n, A, T.n, T.d)
                gs = (1, B)
                args = (vw, tw, y1, y2)
                side_t, side_v, src_t, src_v = 1, 0, V, T
                dwx, dwy = dvw, dtw
            else:
                dims = (A, T.n, B, V.n, T.d)
                gs = (B, 1)
                args = (tw, vw, y1, y2)
                side_t, side_v, src_t, src_v = 0, 1, V, T
                dwx, dwy = dtw, dvw
            if need_t:
                maxsim2_bwd(side_t, src_t, *args, g, gs[0], gs[1], 0.5, *dims, dtn)
            if need_v:
                maxsim2_bwd(side_v, src_v, *args, g, gs[0], gs[1], 0.5, *dims, dvn)
            if dwx is not None or dwy is not None:
                maxsim2_bwd_w(p1, p2, g, gs[0], gs[1], 0.5, *dims[:4], dwx, dwy)
            dtext = T.backward(dtn) if need_t else None
            dvideo = V.backward(dvn) if need_v else None
            ctx.T = ctx.V = None
            return dtext, dvideo, dtw, dvw, None, None, None, None, None
        # direction 1: X = text, Y = video, dH[rx=a, ry=b] = 0.5 g[a,b]
        if need_t:
            vs, vld = V.bwd_source(prec)
            _call("nr_maxsim_bwd_x", prec, _p(vs), vld, _p(tw), _p(tm), _p(vm), _p(y1), _p(g), B, 1, 0.5,
                  A, T.n, B, V.n, T.d, _p(dtn), st)
            _call("nr_maxsim_bwd_y", prec, _p(vs), vld, _p(vw), _p(vm), _p(tm), _p(y2), _p(g), 1, B, 0.5,
                  B, V.n, A, T.n, T.d, _p(dtn), st)
        if need_v:
            ts, tld = T.bwd_source(prec)
            _call("nr_maxsim_bwd_y", prec, _p(ts), tld, _p(tw), _p(tm), _p(vm), _p(y1), _p(g), B, 1, 0.5,
                  A, T.n, B, V.n, T.d, _p(dvn), st)
            _call("nr_maxsim_bwd_x", prec, _p(ts), tld, _p(vw), _p(vm), _p(tm), _p(y2), _p(g), 1, B, 0.5,
                  B, V.n, A, T.n, T.d, _p(dvn), st)
        if need_tw:
            _call("nr_maxsim_bwd_w", _p(p1), _p(g), B, 1, 0.5, A, T.n, B, _p(dtw), st)
        if need_vw:
            _call("nr_maxsim_bwd_w", _p(p2), _p(g), 1, B, 0.5, B, V.n, A, _p(dvw), st)
        dtext = T.backward(dtn) if need_t else None
        dvideo = V.backward(dvn) if need_v else None
        ctx.T = ctx.V = None
        return dtext, dvideo, dtw, dvw, None, None, None, None, None


def maxsim(text_feat, video_feat, tw, vw, text_mask, video_mask, precision="fp32", bwd_precision=None,
           normalize=True):
    prec = PRECISIONS[precision]
    bprec = PRECISIONS[bwd_precision or precision]
    return MaxSimFunction.apply(text_feat, video_feat, tw, vw, text_mask, video_mask, prec, bprec, normalize)


# ------------------------------------------------------------------------------------------------
# token-weight MLP logits: Linear(D,2D) - ReLU - Linear(2D,1)  (reference modeling.py:148-153)
# ------------------------------------------------------------------------------------------------
class _tf32:
    """Scoped TF32 for the library GEMMs of the weight MLP (both passes), without touching global state."""

    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        if self.on:
            torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *a):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


MLP_FP32, MLP_TF32, MLP_BF16 = 0, 1, 2
# bf16 mode: the MLP GEMMs run on this repo's tcgen05 kernel (csrc/gemm_tc.cu); tests flip this to compare against the
# library GEMMs
USE_OWN_GEMM = os.environ.get("NR_OWN_GEMM", "1") != "0"


def mlp_mode(v):
    """Arithmetic of the weight-MLP GEMMs: False / "fp32" -> exact fp32, True / "tf32" -> TF32 tensor cores,
    "bf16" -> bf16 operands (inputs, W1, hidden activations and their gradients stored in bf16, fp32 accumulation
    and fp32 parameter gradients)."""
    if isinstance(v, str):
        return {"fp32": MLP_FP32, "tf32": MLP_TF32, "bf16": MLP_BF16}[v]
    return int(v)


class TokenWeightsFunction(torch.autograd.Function):
    """softmax over tokens of the masked weight-MLP logits (reference modeling.py:485-492) as ONE autograd node
    for the batch tokens and, optionally, the memory-bank tokens of the same modality:
        forward : first layer = library GEMM with bias+ReLU epilogue into one hidden buffer, then
                  nr_token_weights_fwd (second layer + mask + softmax, one pass over h);
        backward: nr_token_weights_bwd (softmax + second layer + ReLU backward, one pass over h) and the dW1 / dx
                  library GEMMs; bank tokens take no dx.
    Replaces ~10 ATen launches per evaluation and the parameter-gradient accumulation of two separate nodes."""

    @staticmethod
    def forward(ctx, xa, ma, xb, mb, w1, b1, w2, b2, mode):
        _req_cuda(xa, xb, w1, b1, w2, b2)
        Ra, N, D = xa.shape
        Rb = xb.shape[0] if xb is not None else 0
        if Rb and tuple(xb.shape[1:]) != (N, D):
            raise RuntimeError("token_weights: batch and bank tokens differ in shape")
        H = w1.shape[0]
        Ta, Tb = Ra * N, Rb * N
        lp = mode == MLP_BF16
        own = lp and USE_OWN_GEMM and D % 8 == 0 and H % 16 == 0 and N <= 128
        ma, mb = _mask(ma), (_mask(mb) if Rb else None)
        w = torch.empty(Ra + Rb, N, dtype=torch.float32, device=xa.device)
        w2c, b2c = _f32c(w2).reshape(-1), _f32c(b2).reshape(-1)
        ctx.mode, ctx.dims, ctx.own = mode, (Ra, Rb, N, D, H), own
        ctx.xshape = xa.shape
        if own:
            # libnrhead.so end to end: bf16 operand copies, ONE tcgen05 GEMM over batch + bank tokens with bias, ReLU
            # and the second layer's dot product in its epilogue, then the masked softmax over tokens
            st = _stream()
            dev = xa.device
            xbf = torch.empty(Ta + Tb, D, dtype=torch.bfloat16, device=dev)
            w1bf = torch.empty(H, D, dtype=torch.bfloat16, device=dev)
            _call("nr_cast_bf16", _p(_f32c(xa)), _p(xbf), Ta * D, st)
            if Rb:
                _call("nr_cast_bf16", _p(_f32c(xb)), _p(xbf[Ta:]), Tb * D, st)
            _call("nr_cast_bf16", _p(_f32c(w1)), _p(w1bf), H * D, st)
            keep = any(ctx.needs_input_grad)
            h = torch.empty(Ta + Tb, H, dtype=torch.bfloat16, device=dev) if keep else None
            logits = torch.zeros(Ra + Rb, N, dtype=torch.float32, device=dev)
            _call("nr_mlp_fwd", _p(xbf), Ta + Tb, D, _p(w1bf), H, _p(_f32c(b1)), _p(w2c), _p(h), _p(logits), st)
            _call("nr_token_softmax", _p(logits), _p(b2c), _p(ma), _p(mb), Ra, Ra + Rb, N, _p(w), st)
            ctx.save_for_backward(xbf, None, h, w, w1bf, w2c)
            return (w[:Ra], w[Ra:]) if Rb else (w[:Ra], None)
        cdt = torch.bfloat16 if lp else torch.float32
        xa2 = _f32c(xa).reshape(Ta, D)
        xb2 = _f32c(xb).reshape(Tb, D) if Rb else None
        w1c, b1c = w1, b1
        if lp:                                  # operand copies; the GEMMs accumulate in fp32
            xa2 = xa2.to(cdt)
            xb2 = xb2.to(cdt) if Rb else None
            w1c, b1c = w1.to(cdt), b1.to(cdt)
        h = torch.empty(Ta + Tb, H, dtype=cdt, device=xa.device)
        with _tf32(mode == MLP_TF32):
            torch._addmm_activation(b1c, xa2, w1c.t(), use_gelu=False, out=h[:Ta])     # bias + ReLU in the epilogue
            if Rb:
                torch._addmm_activation(b1c, xb2, w1c.t(), use_gelu=False, out=h[Ta:])
        _call("nr_token_weights_fwd", _p(h), int(lp), _p(w2c), _p(b2c), _p(ma), _p(mb), Ra, Ra + Rb, N, H, _p(w),
              _stream())
        ctx.save_for_backward(xa2, xb2, h, w, w1c, w2c)
        wa = w[:Ra]
        if Rb:
            return wa, w[Ra:]
        return wa, None

    @staticmethod
    def backward(ctx, dwa, dwb):
        xa2, xb2, h, w, w1c, w2c = ctx.saved_tensors
        Ra, Rb, N, D, H = ctx.dims
        Ta, Tb = Ra * N, Rb * N
        lp = ctx.mode == MLP_BF16
        need = ctx.needs_input_grad
        dwa = _f32c(dwa) if dwa is not None else None
        dwb = _f32c(dwb) if (dwb is not None and Rb) else None
        dh = torch.empty_like(h)
        nch = _lib.load().nr_mlp_chunks(Ta + Tb)
        partials = torch.empty(2 * H + 1, nch, dtype=torch.float32, device=h.device)
        sums = torch.empty(2 * H + 1, dtype=torch.float32, device=h.device)
        dw1 = dx = None
        if ctx.own:
            if need[4]:
                dw1 = torch.zeros(H, D, dtype=torch.float32, device=h.device)
            if need[0]:
                dx = torch.zeros(Ta, D, dtype=torch.float32, device=h.device)
        st = _stream()
        _call("nr_token_weights_bwd", _p(h), int(lp), _p(w), _p(dwa), _p(dwb), Ra, Ra + Rb, N, _p(w2c), H, _p(dh),
              _p(partials), st)
        with ForkJoin(1) as fj:
            with fj.on(0):                        # bias / second-layer gradients next to the GEMMs
                _call("nr_vec_sums", _p(partials), 2 * H + 1, nch, None, _p(sums), _stream())
            if ctx.own:
                # dW1 = dh^T x over batch + bank tokens (operands as stored, split-K), dx = dh W1 for the batch tokens
                if need[4]:
                    _call("nr_mlp_bwd_dw1", _p(dh), Ta + Tb, H, _p(xa2), D, _p(dw1), _stream())
                if need[0]:
                    _call("nr_mlp_bwd_dx", _p(dh), Ta, H, _p(w1c), D, _p(dx), 1, _stream())
                    dx = dx.reshape(ctx.xshape)
            else:
                with _tf32(ctx.mode == MLP_TF32):
                    if need[4]:
                        dw1 = _dw1_splitk(dh[:Ta], xa2)
                        if Rb:
                            dw1 += _dw1_splitk(dh[Ta:], xb2)
                    if need[0]:
                        dx = (torch.mm(dh[:Ta], w1c, out_dtype=torch.float32) if lp else dh[:Ta] @ w1c).reshape(ctx.xshape)
        db1 = sums[:H] if need[5] else None
        dw2 = sums[H:2 * H].reshape(1, H) if need[6] else None
        db2 = sums[2 * H:] if need[7] else None
        return dx, None, None, None, dw1, db1, dw2, db2, None


def _cast_multi(pairs):
    """[(fp32 source, bf16 destination view)] -> one nr_cast_bf16_multi launch."""
    n = len(pairs)
    srcs = (ctypes.c_void_p * n)(*[p_[0].data_ptr() for p_ in pairs])
    dsts = (ctypes.c_void_p * n)(*[p_[1].data_ptr() for p_ in pairs])
    sizes = (ctypes.c_int64 * n)(*[p_[0].numel() for p_ in pairs])
    _call("nr_cast_bf16_multi", ctypes.cast(srcs, ctypes.c_void_p), ctypes.cast(dsts, ctypes.c_void_p),
          ctypes.cast(sizes, ctypes.c_void_p), n, _stream())


class TokenWeightsPairFunction(torch.autograd.Function):
    """The text AND the video token-weight MLP of a head step (batch + bank tokens each) as ONE autograd node on the
    tcgen05 GEMM (bf16 mode): one operand-copy launch, one forward GEMM launch over both modalities, two softmax
    launches; backward: the hidden-layer pass per modality, then ONE GEMM launch for dW1 and dx of both.  A persistent
    GEMM grid per modality would serialise on shared memory and pay its ramp-up four times per step."""

    @staticmethod
    def forward(ctx, xt, mt, xtb, mtb, xv, mv, xvb, mvb, w1t, b1t, w2t, b2t, w1v, b1v, w2v, b2v, bank_bf16=None):
        """bank_bf16 = (text buffer, video buffer) from bank.BankRing.mlp_operands(): [Ta + Tb, D] bf16 whose tail
        already holds the bank rows (persistent across steps); only the batch tokens are cast, into its head."""
        _req_cuda(xt, xtb, xv, xvb, w1t, w1v)
        dev = xt.device
        D, H = xt.shape[2], w1t.shape[0]
        sides, casts = [], []
        for i, (xa, ma, xb, mb, w1, b1, w2, b2) in enumerate(((xt, mt, xtb, mtb, w1t, b1t, w2t, b2t),
                                                              (xv, mv, xvb, mvb, w1v, b1v, w2v, b2v))):
            Ra, N = xa.shape[0], xa.shape[1]
            Rb = xb.shape[0] if xb is not None else 0
            if Rb and tuple(xb.shape[1:]) != (N, D):
                raise RuntimeError("token_weights: batch and bank tokens differ in shape")
            Ta, Tb = Ra * N, Rb * N
            w1bf = torch.empty(H, D, dtype=torch.bfloat16, device=dev)
            if bank_bf16 is not None:
                xbf = bank_bf16[i]
                if tuple(xbf.shape) != (Ta + Tb, D) or xbf.dtype != torch.bfloat16:
                    raise RuntimeError("token_weights: persistent bank operand does not match batch + bank tokens")
                casts += [(_f32c(xa), xbf), (_f32c(w1), w1bf)]
            else:
                xbf = torch.empty(Ta + Tb, D, dtype=torch.bfloat16, device=dev)
                casts += [(_f32c(xa), xbf), (_f32c(w1), w1bf)]
                if Rb:
                    casts.append((_f32c(xb), xbf[Ta:]))
            sides.append(dict(Ra=Ra, Rb=Rb, N=N, Ta=Ta, Tb=Tb, xbf=xbf, w1bf=w1bf, ma=_mask(ma), mb=_mask(mb) if Rb else None,
                              b1=_f32c(b1), w2=_f32c(w2).reshape(-1), b2=_f32c(b2).reshape(-1), xshape=xa.shape))
        _cast_multi(casts)
        keep = any(ctx.needs_input_grad)
        arr = (_lib.MlpSide * 2)()
        # the second-layer dot products accumulate atomically: ONE zero fill for both modalities' logits
        nlog = [(sd["Ra"] + sd["Rb"]) * sd["N"] for sd in sides]
        zlog = torch.zeros(sum(nlog), dtype=torch.float32, device=dev)
        for i, sd in enumerate(sides):
            sd["h"] = torch.empty(sd["Ta"] + sd["Tb"], H, dtype=torch.bfloat16, device=dev) if keep else None
            sd["logits"] = zlog[sum(nlog[:i]):sum(nlog[:i + 1])].view(sd["Ra"] + sd["Rb"], sd["N"])
            sd["w"] = torch.empty(sd["Ra"] + sd["Rb"], sd["N"], dtype=torch.float32, device=dev)
            a = arr[i]
            a.x_bf16, a.w1_bf16, a.T = sd["xbf"].data_ptr(), sd["w1bf"].data_ptr(), sd["Ta"] + sd["Tb"]
            a.b1, a.w2, a.logits = sd["b1"].data_ptr(), sd["w2"].data_ptr(), sd["logits"].data_ptr()
            a.h_bf16 = sd["h"].data_ptr() if keep else None
        st = _stream()
        # one SM stays free for the single-CTA Sinkhorn kernel that runs next to this GEMM in the head's forward
        _call("nr_mlp_fwd_pair", ctypes.cast(arr, ctypes.c_void_p), 2, D, H, 1, st)
        sm = (_lib.SoftmaxSide * 2)()
        for i, sd in enumerate(sides):
            q = sm[i]
            q.logits, q.b2, q.w = sd["logits"].data_ptr(), sd["b2"].data_ptr(), sd["w"].data_ptr()
            q.mask_a = sd["ma"].data_ptr() if sd["ma"] is not None else None
            q.mask_b = sd["mb"].data_ptr() if sd["mb"] is not None else None
            q.Ra, q.R, q.N = sd["Ra"], sd["Ra"] + sd["Rb"], sd["N"]
        _call("nr_token_softmax_pair", ctypes.cast(sm, ctypes.c_void_p), 2, st)
        ctx.dims = [(sd["Ra"], sd["Rb"], sd["N"], sd["xshape"]) for sd in sides] + [(D, H)]
        # the split-K accumulators of the backward GEMM (dW1, dx of both modalities), zeroed next to the forward
        need = ctx.needs_input_grad
        ctx.zsizes = [H * D if need[8] else 0, sides[0]["Ta"] * D if need[0] else 0,
                      H * D if need[12] else 0, sides[1]["Ta"] * D if need[4] else 0]
        ctx.prezero = prezeroed(sum(ctx.zsizes), dev, 7) if keep else None
        saved = []
        for sd in sides:
            saved += [sd["xbf"], sd["h"], sd["w"], sd["w1bf"], sd["w2"]]
        ctx.save_for_backward(*saved)
        out = []
        for sd in sides:
            out += [sd["w"][:sd["Ra"]], sd["w"][sd["Ra"]:] if sd["Rb"] else None]
        return tuple(out)

    @staticmethod
    def backward(ctx, dwt, dwtb, dwv, dwvb):
        sv = ctx.saved_tensors
        D, H = ctx.dims[2]
        need = ctx.needs_input_grad
        dev = sv[0].device
        f32 = dict(dtype=torch.float32, device=dev)
        arr = (_lib.MlpSide * 2)()
        res = []
        keepalive = []
        if ctx.prezero is not None:
            z, ev_z = ctx.prezero
            torch.cuda.current_stream().wait_event(ev_z)
        else:
            z = torch.zeros(sum(ctx.zsizes), **f32)             # ONE fill for the four accumulators
        zo = [0]

        def take(n, shape):
            t = z[zo[0]:zo[0] + n].view(shape)
            zo[0] += n
            return t

        # every buffer before the fork
        for i, (dwa, dwb, nx, nw1) in enumerate(((dwt, dwtb, need[0], need[8]), (dwv, dwvb, need[4], need[12]))):
            xbf, h, w, w1bf, w2c = sv[5 * i:5 * i + 5]
            Ra, Rb, N, xshape = ctx.dims[i]
            Ta, T = Ra * N, (Ra + Rb) * N
            nch = _lib.load().nr_mlp_chunks(T)
            res.append(dict(xbf=xbf, h=h, w=w, w1bf=w1bf, w2c=w2c, Ra=Ra, Rb=Rb, N=N, Ta=Ta, T=T, xshape=xshape, nx=nx, nw1=nw1,
                            dwa=_f32c(dwa) if dwa is not None else None, dwb=_f32c(dwb) if (dwb is not None and Rb) else None,
                            dh=torch.empty_like(h), partials=torch.empty(2 * H + 1, nch, **f32), nch=nch,
                            sums=torch.empty(2 * H + 1, **f32),
                            dw1=take(H * D, (H, D)) if nw1 else None, dx=take(Ta * D, (Ta, D)) if nx else None))
        def hidden_bwd(r):
            _call("nr_token_weights_bwd", _p(r["h"]), 1, _p(r["w"]), _p(r["dwa"]), _p(r["dwb"]), r["Ra"], r["Ra"] + r["Rb"],
                  r["N"], _p(r["w2c"]), H, _p(r["dh"]), _p(r["partials"]), _stream())

        with ForkJoin(1, offset=4) as fj:
            hidden_bwd(res[0])
            with fj.on(0):
                hidden_bwd(res[1])
        for i, r in enumerate(res):
            a = arr[i]
            a.x_bf16, a.w1_bf16, a.T, a.T_dx = r["xbf"].data_ptr(), r["w1bf"].data_ptr(), r["T"], r["Ta"]
            a.dh_bf16 = r["dh"].data_ptr()
            a.dw1 = r["dw1"].data_ptr() if r["dw1"] is not None else None
            a.dx = r["dx"].data_ptr() if r["dx"] is not None else None
        with ForkJoin(1, offset=4) as fj:
            with fj.on(0):                        # bias / second-layer gradients next to the GEMMs
                for r in res:
                    _call("nr_vec_sums", _p(r["partials"]), 2 * H + 1, r["nch"], None, _p(r["sums"]), _stream())
            _call("nr_mlp_bwd_pair", ctypes.cast(arr, ctypes.c_void_p), 2, D, H, _stream())
            if EVENTS.get("_want_bank_events"):       # last reader of the bank's MLP operand (graph.py: late insert)
                ev_b = torch.cuda.Event()
                ev_b.record()
                EVENTS["mlp_gemm_bwd_done"] = ev_b
        out = []
        for i, r in enumerate(res):
            base = 8 + 4 * i
            s_ = r["sums"]
            out.append((r["dx"].reshape(r["xshape"]) if r["dx"] is not None else None,
                        r["dw1"], s_[:H] if need[base + 1] else None, s_[H:2 * H].reshape(1, H) if need[base + 2] else None,
                        s_[2 * H:] if need[base + 3] else None))
        (dxt, dw1t, db1t, dw2t, db2t), (dxv, dw1v, db1v, dw2v, db2v) = out
        # The sharded captured step reduces the video-feature gradients of the head on a side stream (sharded.py);
        # autograd adds dxv to them right after this node: order that addition behind the collective.  Enqueued AFTER
        # this node's own launches, so it delays nothing of the MLP backward.
        ev = EVENTS.get("video_grad_ready")
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        return (dxt, None, None, None, dxv, None, None, None, dw1t, db1t, dw2t, db2t, dw1v, db1v, dw2v, db2v, None)


def token_weights_pair(text_mlp, video_mlp, text, text_mask, video, video_mask, mode, bank_t=None, bank_mt=None,
                       bank_v=None, bank_mv=None, bank_bf16=None):
    """(tw, tw_bank, vw, vw_bank): both token-weight MLPs of a step.  bf16 mode on the tcgen05 GEMM: one merged node
    (TokenWeightsPairFunction); any other mode: the two per-modality nodes, the video one on a forked stream."""
    pt = text_mlp if isinstance(text_mlp, (tuple, list)) else mlp_params(text_mlp)
    pv = video_mlp if isinstance(video_mlp, (tuple, list)) else mlp_params(video_mlp)
    D, H = text.shape[2], pt[0].shape[0]
    if (mlp_mode(mode) == MLP_BF16 and USE_OWN_GEMM and D % 8 == 0 and H % 16 == 0 and text.shape[1] <= 128
            and video.shape[1] <= 128 and video.shape[2] == D and pv[0].shape[0] == H):
        return TokenWeightsPairFunction.apply(text, text_mask, bank_t, bank_mt, video, video_mask, bank_v, bank_mv, *pt, *pv,
                                              bank_bf16)
    with ForkJoin(1, offset=5) as fj:
        tw, tw_mb = token_weights(pt, text, text_mask, mode, bank_t, bank_mt)
        with fj.on(0):
            vw, vw_mb = token_weights(pv, video, video_mask, mode, bank_v, bank_mv)
            vw.record_stream(fj.main)
    return tw, tw_mb, vw, vw_mb


def _dw1_splitk(dh, x):
    """dh^T @ x for dh [T,H], x [T,D] with T >> H, D: the library runs this [H,D] output as 64 CTAs over the whole
    K = T; a batched product over K-chunks (split-K) fills the GPU, the chunk sum is one small reduction."""
    T = x.shape[0]
    lp = dh.dtype != torch.float32                  # bf16 operands: fp32 accumulation AND fp32 output
    for parts in (8, 6, 4, 3, 2):
        if T % parts == 0 and T // parts >= 1024:
            c = T // parts
            a3, b3 = dh.view(parts, c, dh.shape[1]).transpose(1, 2), x.view(parts, c, x.shape[1])
            return (torch.bmm(a3, b3, out_dtype=torch.float32) if lp else torch.bmm(a3, b3)).sum(0)
    return torch.mm(dh.t(), x, out_dtype=torch.float32) if lp else dh.t() @ x


def token_weights(mlp, feat, mask, mode, bank_feat=None, bank_mask=None):
    """mlp: nn.Sequential(Linear, ReLU, Linear) with the reference's parameter names, or its 4 parameters
    (W1, b1, W2, b2) as a tuple; mode: see mlp_mode().  Returns the token weights of `feat` [R,N] and of
    `bank_feat` (or None)."""
    ps = mlp if isinstance(mlp, (tuple, list)) else mlp_params(mlp)
    return TokenWeightsFunction.apply(feat, mask, bank_feat, bank_mask, *ps, mlp_mode(mode))


def mlp_params(mlp):
    return (mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias)


# ------------------------------------------------------------------------------------------------
# centrality weights (reference modeling.py:403-430)
# ------------------------------------------------------------------------------------------------
class CentralityWeightsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, gfeat, cs):
        _req_cuda(feat, gfeat)
        P = Prepared(feat, colsum=True)
        g = _f32c(gfeat).reshape(-1, gfeat.shape[-1])
        if g.shape[0] != feat.shape[0]:
            # the reference broadcasts [B,G>1] weights into a shape error
            raise RuntimeError("centrality weights need exactly one global token per sample "
                               f"(got {tuple(gfeat.shape)})")
        B, d = g.shape
        dev = g.device
        mean_vec = torch.empty(d, dtype=torch.float32, device=dev)
        gn = torch.empty(B, d, dtype=torch.float32, device=dev)
        ginv = torch.empty(B, dtype=torch.float32, device=dev)
        w = torch.empty(B, dtype=torch.float32, device=dev)
        _call("nr_centrality_fwd", _p(P.partials), P.partials.shape[0], P.rows, _p(g), B, d, float(cs), _p(mean_vec),
              _p(gn), _p(ginv), _p(w), _stream(), launches=2)
        ctx.P, ctx.cs, ctx.gshape = P, float(cs), gfeat.shape
        ctx.save_for_backward(mean_vec, gn, ginv, w)
        return w

    @staticmethod
    def backward(ctx, dw):
        mean_vec, gn, ginv, w = ctx.saved_tensors
        P = ctx.P
        B, d = gn.shape
        dw = _f32c(dw)
        need_f, need_g = ctx.needs_input_grad[:2]
        dg = torch.empty(B, d, dtype=torch.float32, device=gn.device) if need_g else None
        dmean = torch.empty(d, dtype=torch.float32, device=gn.device) if need_f else None
        _call("nr_centrality_bwd", _p(mean_vec), _p(gn), _p(ginv), _p(w), _p(dw), B, d, ctx.cs, P.rows, _p(dg), 0,
              _p(dmean), _stream(), launches=2)
        dfeat = P.backward(None, add_vec=dmean) if need_f else None
        ctx.P = None
        return dfeat, (dg.reshape(ctx.gshape) if need_g else None), None


def centrality_weights(feat, gfeat, cs):
    return CentralityWeightsFunction.apply(feat, gfeat, cs)


# ------------------------------------------------------------------------------------------------
# Sinkhorn duals (no grad; reference until_module.py:222-251)
# ------------------------------------------------------------------------------------------------
def sinkhorn_workspace(B, dev):
    n = _lib.load().nr_sinkhorn_workspace_bytes(B)
    return torch.empty(n, dtype=torch.uint8, device=dev), n


def sinkhorn_duals(G, GT, iters=50):
    """Both directions in one cooperative launch: returns (u1, v1, u2, v2); chain 1 on G, chain 2 on G^T."""
    _req_cuda(G, GT)
    G, GT = _f32c(G.detach()), _f32c(GT.detach())
    B = G.shape[0]
    if G.shape != (B, B) or GT.shape != (B, B):
        raise RuntimeError("sinkhorn_duals: square [B,B] matrices required")
    duals = torch.empty(4, B, dtype=torch.float32, device=G.device)
    lib = _lib.load()
    nbytes = lib.nr_sinkhorn_workspace_bytes(B)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=G.device)
    _call("nr_sinkhorn", _p(G), _p(GT), B, int(iters), _p(duals[0]), _p(duals[1]), _p(duals[2]), _p(duals[3]),
          _p(ws), nbytes, _stream(), launches=2)
    return duals[0], duals[1], duals[2], duals[3]


# ------------------------------------------------------------------------------------------------
# row-block losses
# ------------------------------------------------------------------------------------------------
class RowLossesFunction(torch.autograd.Function):
    """Per-row loss terms of ONE direction over a row block.  Returns the 4 row sums
    {centrality, neighbour, kl, uniform} (un-normalised: divide by B or B^2 outside)."""

    @staticmethod
    def forward(ctx, X, G, cbank, w, logit_scale, sk_u, sk_v, row0, k, tau_nbr, tau_uni, beta, flags):
        _req_cuda(X, G, cbank, w, logit_scale, sk_u, sk_v)
        X = _f32c(X)
        rows, B = X.shape
        dev = X.device
        G = _f32c(G) if G is not None else None
        cbank = _f32c(cbank) if cbank is not None else None
        w = _f32c(w) if w is not None else None
        ls = _f32c(logit_scale).reshape(1) if logit_scale is not None else None
        row_out = torch.zeros(4, rows, dtype=torch.float32, device=dev)
        nbr = torch.empty(rows, max(int(k), 1), dtype=torch.int32, device=dev)
        saved = torch.zeros(rows, NR_NSAVE, dtype=torch.float32, device=dev)
        st = _stream()
        _call("nr_row_losses_fwd", _p(X), B, _p(G), B, _p(cbank), _p(w), _p(sk_u), _p(sk_v), rows, B, int(row0),
              _p(ls), int(k), float(tau_nbr), float(tau_uni), float(beta), int(flags), _p(row_out), _p(nbr),
              _p(saved), st)
        sums = torch.empty(4, dtype=torch.float32, device=dev)
        _call("nr_vec_sums", _p(row_out), 4, rows, None, _p(sums), st)
        ctx.cfg = (int(row0), int(k), float(tau_nbr), float(tau_uni), float(beta), int(flags))
        ctx.save_for_backward(X, G, cbank, w, ls, sk_u, sk_v, nbr, saved)
        ctx.mark_non_differentiable(nbr)
        return sums, nbr

    @staticmethod
    def backward(ctx, dsums, _dnbr):
        X, G, cbank, w, ls, sk_u, sk_v, nbr, saved = ctx.saved_tensors
        row0, k, tau_nbr, tau_uni, beta, flags = ctx.cfg
        rows, B = X.shape
        dev = X.device
        gscale = _f32c(dsums)
        dX = torch.empty_like(X)
        dG = torch.empty_like(G) if (G is not None and ctx.needs_input_grad[1]) else None
        dc = torch.zeros(B, dtype=torch.float32, device=dev) if (cbank is not None and ctx.needs_input_grad[2]) else None
        dw = torch.zeros(rows, dtype=torch.float32, device=dev) if (w is not None and ctx.needs_input_grad[3]) else None
        dls = torch.zeros(1, dtype=torch.float32, device=dev) if (ls is not None and ctx.needs_input_grad[4]) else None
        _call("nr_row_losses_bwd", _p(X), B, _p(G), B, _p(cbank), _p(w), _p(sk_u), _p(sk_v), rows, B, row0, _p(ls), k,
              tau_nbr, tau_uni, beta, flags, _p(nbr), _p(saved), _p(gscale), _p(dX), B, _p(dG), B, _p(dc), _p(dw),
              _p(dls), _stream())
        if dls is not None:
            dls = dls.reshape(())
        return dX, dG, dc, dw, dls, None, None, None, None, None, None, None, None


def row_losses(X, G=None, cbank=None, w=None, logit_scale=None, sk_u=None, sk_v=None, row0=0, k=1, tau_nbr=1.0,
               tau_uni=1.0, beta=0.0, flags=0):
    if logit_scale is not None and logit_scale.dim() != 0 and logit_scale.numel() != 1:
        raise RuntimeError("logit_scale must be a scalar tensor")
    return RowLossesFunction.apply(X, G, cbank, w, logit_scale, sk_u, sk_v, row0, k, tau_nbr, tau_uni, beta, flags)


class RowMeanFunction(torch.autograd.Function):
    """memory_bank_matrix.sum(-1) / size(-1)  (reference until_module.py:181)."""

    @staticmethod
    def forward(ctx, X):
        _req_cuda(X)
        X = _f32c(X)
        out = torch.empty(X.shape[0], dtype=torch.float32, device=X.device)
        _call("nr_row_mean", _p(X), X.shape[1], X.shape[0], X.shape[1], _p(out), _stream())
        ctx.shape = X.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        return (dout / ctx.shape[1]).unsqueeze(1).expand(ctx.shape)


def row_mean(X):
    return RowMeanFunction.apply(X)


# ------------------------------------------------------------------------------------------------
# memory bank FIFO and evaluation ranking (no autograd)
# ------------------------------------------------------------------------------------------------
def fifo_update(new, old, capacity):
    """cat(new, old)[:capacity] along dim 0 in one copy kernel (reference modeling.py:235-249)."""
    _req_cuda(new, old)
    new, old = new.contiguous(), old.contiguous()
    if new.dtype != old.dtype or new.shape[1:] != old.shape[1:]:
        raise RuntimeError("fifo_update: new/old rows differ in dtype or shape")
    out = torch.empty((capacity,) + tuple(new.shape[1:]), dtype=new.dtype, device=new.device)
    row_bytes = new[0].numel() * new.element_size() if new.shape[0] else old[0].numel() * old.element_size()
    _call("nr_fifo_update", _p(new), new.shape[0], _p(old), old.shape[0], _p(out), capacity, row_bytes, _stream())
    return out


def rank_counts(S, diag=None, diag_col0=0, gt=None, eq=None):
    """(#greater, #equal) per query row against its positive's score (reference metrics.py:58-66)."""
    _req_cuda(S)
    S = _f32c(S)
    Q, N = S.shape
    if gt is None:
        gt = torch.zeros(Q, dtype=torch.int32, device=S.device)
        eq = torch.zeros(Q, dtype=torch.int32, device=S.device)
    _call("nr_rank_count", _p(S), N, Q, N, _p(diag), int(diag_col0), _p(gt), _p(eq), _stream())
    return gt, eq


class FusedRanker:
    """Evaluation ranks straight from the tensor-core accumulator (nr_maxsim2_rank): the [Q, N] similarity block is
    never written.  Texts are the pairs text0 .. text0+Q-1, videos the pairs video0 .. video0+N-1 of a square test
    set; `diagonal(total)` returns the positives' scores this block holds (zeros elsewhere: shards add theirs with
    an all-reduce), `counts(diag)` the (#greater, #equal) vectors of both retrieval directions — what
    rank_counts gives on the materialised block and on its transpose (reference utils/metrics.py:58-66)."""

    def __init__(self, text_feat, video_feat, tw, vw, text_mask, video_mask, precision="bf16", text0=0, video0=0):
        _req_cuda(text_feat, video_feat, tw, vw, text_mask, video_mask)
        prec = PRECISIONS[precision]
        if prec not in TC_PRECISIONS:
            raise RuntimeError("FusedRanker: the ranks come from the tensor-core kernel (precision 'bf16' or 'bf16x3')")
        x3 = prec == NR_PREC_BF16X3
        swap = _fused_orientation(text_feat.shape[1], video_feat.shape[1], text_feat.shape[2] * (3 if x3 else 1))
        if swap is None:
            raise RuntimeError("FusedRanker: token counts without a fused tensor-core instantiation")
        self.swap, self.flags = swap, 1 if x3 else 0
        self.tw, self.vw = _f32c(tw), _f32c(vw)
        T = Prepared(text_feat, bf16=True, f32=False, mask=_mask(text_mask),
                     split=(ROLE_Y if swap else ROLE_X) if x3 else 0)
        V = Prepared(video_feat, bf16=True, f32=False, mask=_mask(video_mask),
                     split=(ROLE_X if swap else ROLE_Y) if x3 else 0)
        self.T, self.V, self.text0, self.video0 = T, V, int(text0), int(video0)

    def _launch(self, mode, diag, cnt):
        T, V = self.T, self.V
        q = _lib.MaxSim2RankProblem()
        X, Y, wx, wy, gx0, gy0 = ((V, T, self.vw, self.tw, self.video0, self.text0) if self.swap else
                                  (T, V, self.tw, self.vw, self.text0, self.video0))
        if diag.dtype != torch.float32 or not diag.is_contiguous() or diag.numel() < max(gx0 + X.r, gy0 + Y.r):
            raise RuntimeError("FusedRanker: diag must be a contiguous fp32 vector covering every pair id of the block")
        q.x_bf16, q.y_bf16, q.wx, q.wy = X.xn_bf16.data_ptr(), Y.xn_bf16.data_ptr(), wx.data_ptr(), wy.data_ptr()
        q.Rx, q.Ry, q.alpha, q.gx0, q.gy0, q.diag = X.r, Y.r, 0.5, gx0, gy0, diag.data_ptr()
        if cnt is not None:
            t, v = (cnt[0], cnt[1]), (cnt[2], cnt[3])
            cx, cy = (v, t) if self.swap else (t, v)
            q.gt_x, q.eq_x, q.gt_y, q.eq_y = cx[0].data_ptr(), cx[1].data_ptr(), cy[0].data_ptr(), cy[1].data_ptr()
        ws = _tile_workspace(X.device)
        _call("nr_maxsim2_rank", ctypes.byref(q), mode, X.n, Y.n, X.kd, _p(ws), self.flags, _stream())

    def diagonal(self, total):
        """fp32 [total]: S[positive] at the pair ids whose text AND video are in this block, 0 elsewhere."""
        diag = torch.zeros(int(total), dtype=torch.float32, device=self.T.device)
        self._launch(1, diag, None)
        return diag

    def counts(self, diag):
        """int32 (gt_t [Q], eq_t [Q], gt_v [N], eq_v [N]): per text row over this block's videos, per video over this
        block's texts."""
        dev = self.T.device
        cnt = (torch.zeros(self.T.r, dtype=torch.int32, device=dev), torch.zeros(self.T.r, dtype=torch.int32, device=dev),
               torch.zeros(self.V.r, dtype=torch.int32, device=dev), torch.zeros(self.V.r, dtype=torch.int32, device=dev))
        self._launch(2, diag, cnt)
        return cnt


def topk_rows(S, k, col_offset=0):
    _req_cuda(S)
    S = _f32c(S)
    Q, N = S.shape
    vals = torch.empty(Q, k, dtype=torch.float32, device=S.device)
    idx = torch.empty(Q, k, dtype=torch.int32, device=S.device)
    _call("nr_topk_rows", _p(S), N, Q, N, int(k), int(col_offset), _p(vals), _p(idx), _stream())
    return vals, idx


def topk_merge(vals, idx):
    """[W,Q,k] per-shard lists -> global [Q,k]."""
    _req_cuda(vals, idx)
    vals, idx = _f32c(vals), idx.to(torch.int32).contiguous()
    W, Q, k = vals.shape
    ov = torch.empty(Q, k, dtype=torch.float32, device=vals.device)
    oi = torch.empty(Q, k, dtype=torch.int32, device=vals.device)
    _call("nr_topk_merge", _p(vals), _p(idx), W, Q, k, _p(ov), _p(oi), _stream())
    return ov, oi


def rank_counts_target(S, target, diag=None, col_offset=0, gt=None, eq_before=None, want_valid=True):
    """Multi-sentence text->video ranks (reference metrics.py:81-122): per caption row q of S [Q,N] with its
    video in global column target[q], (#greater-or-NaN, #equal in a lower column, valid) — rank = sum of the
    first two.  gt/eq_before accumulate when passed in (column shards)."""
    _req_cuda(S, target, diag)
    S = _f32c(S)
    Q, N = S.shape
    target = target.to(torch.int32).contiguous()
    if target.numel() != Q:
        raise ValueError(f"rank_counts_target: {target.numel()} targets for {Q} rows")
    if gt is None:
        gt = torch.zeros(Q, dtype=torch.int32, device=S.device)
        eq_before = torch.zeros(Q, dtype=torch.int32, device=S.device)
    valid = torch.empty(Q, dtype=torch.int32, device=S.device) if want_valid else None
    _call("nr_rank_count_target", _p(S), N, Q, N, _p(target), _p(diag), int(col_offset), _p(gt), _p(eq_before),
          _p(valid), _stream())
    return gt, eq_before, valid


def group_max_t(S, group_start):
    """out[j, i] = max over caption rows [group_start[i], group_start[i+1]) of S[t, j], NaN read as -inf
    (reference metrics.py:124-145).  S [T,V] CUDA, group_start int32 [G+1] CUDA -> [V, G] f32."""
    _req_cuda(S, group_start)
    S = _f32c(S)
    T, V = S.shape
    group_start = group_start.to(torch.int32).contiguous()
    G = group_start.numel() - 1
    out = torch.empty(V, G, dtype=torch.float32, device=S.device)
    _call("nr_group_max_t", _p(S), V, T, V, _p(group_start), G, _p(out), G, _stream())
    return out
