"""Persistent prepared memory bank (reference NeighborRetr/models/modeling.py:175-184, 222-249 and
NeighborRetr/utils/memory_bank.py:206-211).

The reference keeps the bank as five plain tensors and rebuilds them every step with ``cat(new, old)[:capacity]``; the
head then re-normalises, re-casts and re-transposes all M rows although only the newest B changed.  ``BankRing`` keeps
everything the head reads of the bank in place, as a ring of M sample slots:

    raw fp32 rows + int64 masks + dataset indices      (what the public ``mb_*`` attributes show, rotated on demand)
    normalised bf16 operand copy (plain, or split for bf16x3), masked tokens zeroed    -> forward contraction
    its transposed copy                                                                  -> backward contraction
    raw bf16 rows behind a scratch area for the batch tokens                             -> weight-MLP GEMM operand

and a step writes only its new samples (csrc/prep.cu: nr_bank_advance + nr_bank_insert, ~10 bytes per new element
instead of ~26 per BANK element).  The ring position is an int32 in device memory, so the insert replays inside the
step's CUDA graph; reference row i (newest first) is slot (head + i) mod M, which is how ``export()`` materialises
the reference's tensors when someone reads ``model.mb_feat_t``.  The similarity columns the head computes against
the bank are only ever averaged over the bank (until_module.py:181), so their order is free.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from .ops import Prepared, _call, _f32c, _mask, _p, _stream

NAMES = ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v")


class BankRing:
    def __init__(self, ind, feat_t, feat_v, mask_t, mask_v, x3=False, mlp_bf16=True, batch_rows=0):
        """Tensors in the reference's order (row 0 = newest).  batch_rows: largest per-step batch (samples) whose
        tokens share the MLP operand buffer with the bank rows."""
        ops._req_cuda(feat_t, feat_v)
        dev = feat_t.device
        self.device = dev
        self.M, self.nt, self.d = feat_t.shape
        self.nv = feat_v.shape[1]
        self.x3 = bool(x3)
        self.dtypes = {n: t.dtype for n, t in zip(NAMES, (ind, feat_t, feat_v, mask_t, mask_v))}
        M, nt, nv, d = self.M, self.nt, self.nv, self.d
        kd = 3 * d if x3 else d
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.ind = torch.empty(M, dtype=torch.int64, device=dev)
        self.feat_t = torch.empty(M, nt, d, dtype=torch.float32, device=dev)
        self.feat_v = torch.empty(M, nv, d, dtype=torch.float32, device=dev)
        self.mask_t = torch.empty(M, nt, dtype=torch.int64, device=dev)
        self.mask_v = torch.empty(M, nv, dtype=torch.int64, device=dev)
        self.head_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.head = 0                                            # host mirror (every insert is a known shift)
        self.ld_t, self.ld_v = (M * nt + 7) // 8 * 8, (M * nv + 7) // 8 * 8
        self.xn_t = torch.empty(M, nt, kd, **bf); self.xn_v = torch.empty(M, nv, kd, **bf)
        self.xnT_t = torch.zeros(kd, self.ld_t, **bf); self.xnT_v = torch.zeros(kd, self.ld_v, **bf)
        self.batch_rows = int(batch_rows)
        if mlp_bf16:
            self.mlp_t = torch.empty((self.batch_rows + M) * nt, d, **bf)
            self.mlp_v = torch.empty((self.batch_rows + M) * nv, d, **bf)
        else:
            self.mlp_t = self.mlp_v = None
        rx, ry = (ops.ROLE_X, ops.ROLE_Y) if x3 else (0, 0)
        self.roles = (rx, ry)
        self.MT = self._view(self.xn_t, self.xnT_t, self.mask_t, nt, rx)
        self.MV = self._view(self.xn_v, self.xnT_v, self.mask_v, nv, ry)
        self.version = 0                  # bumped by every load / insert (cached exports, pending-assignment checks)
        self._export = None
        self.load(ind, feat_t, feat_v, mask_t, mask_v)

    def _view(self, xn, xnT, mask, n, role):
        v = Prepared.__new__(Prepared)
        v.r, v.n, v.d, v.rows = self.M, n, self.d, self.M * n
        v.xn, v.xn_bf16, v.xnT_bf16 = None, xn, xnT
        v.inv_norm = v.partials = None
        v._x, v._t_ready, v._parent, v._lo = None, True, None, 0
        v.mask, v.split, v.device = mask, role, self.device
        return v

    def matches(self, feat_t, feat_v, x3, mlp_bf16, batch_rows):
        return (tuple(feat_t.shape) == (self.M, self.nt, self.d) and tuple(feat_v.shape) == (self.M, self.nv, self.d)
                and self.x3 == bool(x3) and (self.mlp_t is not None) == bool(mlp_bf16) and batch_rows <= self.batch_rows
                and feat_t.device == self.device)

    # ---- (re)build from tensors in reference order: slot i = row i, head = 0 ------------------------------------
    def load(self, ind, feat_t, feat_v, mask_t, mask_v):
        with torch.no_grad():
            self.ind.copy_(ind.reshape(-1)); self.feat_t.copy_(feat_t); self.feat_v.copy_(feat_v)
            self.mask_t.copy_(mask_t); self.mask_v.copy_(mask_v)
            self.head_dev.zero_()
        self.head = 0
        st = _stream()
        for feat, mask, xn, xnT, n, role, mlp in ((self.feat_t, self.mask_t, self.xn_t, self.xnT_t, self.nt, self.roles[0], self.mlp_t),
                                                  (self.feat_v, self.mask_v, self.xn_v, self.xnT_v, self.nv, self.roles[1], self.mlp_v)):
            rows = self.M * n
            if role:
                _call("nr_prep_tokens_split", _p(feat), rows, self.d, None, _p(xn), role, None, None, _p(mask), st)
            else:
                _call("nr_prep_tokens", _p(feat), rows, self.d, None, _p(xn), None, None, _p(mask), st)
            _call("nr_transpose_tokens_bf16", _p(xn), rows, xn.shape[2], _p(xnT), xnT.shape[1], st)
            if mlp is not None:
                _call("nr_cast_bf16", _p(feat), _p(mlp[self.batch_rows * n:]), rows * self.d, st)
        self.version += 1
        self._export = None

    # ---- one step: the newest rows replace the oldest ------------------------------------------------------------
    def insert(self, ind, feat_t, feat_v, mask_t, mask_v, phase="all"):
        """cat(new, bank)[:M] of the reference (modeling.py:235-249), in place.  Must be enqueued after every reader of
        the step (forward AND backward contractions, the MLP GEMMs).
        phase: "all", or the two halves of a split insert — "early": ring position, indices, raw rows, masks and the
        contraction operands (their last reader is the backward contraction); "late": the MLP GEMM's bf16 operand
        (its last reader is the backward GEMM).  A split insert is "early" then "late", exactly once each."""
        if phase not in ("all", "early", "late"):
            raise ValueError(f"BankRing.insert: unknown phase {phase!r}")
        B = feat_t.shape[0]
        n_new = min(B, self.M)
        st = _stream()
        if phase != "late":
            ind = ind.reshape(-1).to(torch.int64).contiguous()
            _call("nr_bank_advance", _p(self.head_dev), n_new, self.M, _p(ind), _p(self.ind), st)
        elif self.mlp_t is None:
            return
        arr = (_lib.BankSide * 2)()
        keep = []
        for i, (new, mask, feat, rmask, xn, xnT, ld, n, role, mlp) in enumerate((
                (feat_t, mask_t, self.feat_t, self.mask_t, self.xn_t, self.xnT_t, self.ld_t, self.nt, self.roles[0], self.mlp_t),
                (feat_v, mask_v, self.feat_v, self.mask_v, self.xn_v, self.xnT_v, self.ld_v, self.nv, self.roles[1], self.mlp_v))):
            raw = mlp[self.batch_rows * n:] if mlp is not None else None
            newc, maskc = _f32c(new.detach()), _mask(mask)
            keep += [newc, maskc]
            a = arr[i]
            a.new_feat, a.new_mask, a.N = newc.data_ptr(), maskc.data_ptr(), n
            a.split_role, a.ld = role, ld
            if phase != "late":
                a.ring_feat, a.ring_mask = feat.data_ptr(), rmask.data_ptr()
                a.ring_xn_bf16, a.ring_xnT_bf16 = xn.data_ptr(), xnT.data_ptr()
            if phase != "early":
                a.ring_raw_bf16 = raw.data_ptr() if raw is not None else None
        _call("nr_bank_insert_pair", ctypes.cast(arr, ctypes.c_void_p), 2, n_new, self.d, self.M, _p(self.head_dev), st)
        if phase == "late":
            return                                # the ring position moved with the early half
        self.head = (self.head - n_new) % self.M
        self.version += 1
        self._export = None

    def note_replay(self, n_new):
        """A CUDA-graph replay ran insert() on the device: advance the host mirror."""
        self.head = (self.head - min(n_new, self.M)) % self.M
        self.version += 1
        self._export = None

    # ---- the reference's tensors (newest first), materialised on demand ------------------------------------------
    def export(self):
        if self._export is None:
            h = self.head
            rot = lambda t, name: torch.roll(t, shifts=-h, dims=0).to(self.dtypes[name])
            self._export = {"mb_ind": rot(self.ind, "mb_ind"), "mb_feat_t": rot(self.feat_t, "mb_feat_t"),
                            "mb_feat_v": rot(self.feat_v, "mb_feat_v"), "mb_mask_t": rot(self.mask_t, "mb_mask_t"),
                            "mb_mask_v": rot(self.mask_v, "mb_mask_v")}
        return self._export

    def mlp_operands(self, b_rows):
        """(text buffer, video buffer): [b_rows*N + M*N, D] bf16 views whose tail is the bank and whose head is scratch
        for this step's batch tokens."""
        if self.mlp_t is None or b_rows > self.batch_rows:
            return None
        o = self.batch_rows - b_rows
        return self.mlp_t[o * self.nt:], self.mlp_v[o * self.nv:]
