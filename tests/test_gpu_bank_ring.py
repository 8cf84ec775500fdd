"""Persistent prepared memory bank (neighborretr_b200/bank.py, csrc/prep.cu nr_bank_advance / nr_bank_insert):
the ring must show exactly the reference's bank — cat(new, old)[:capacity], newest first (reference
NeighborRetr/models/modeling.py:235-249) — through the public mb_* attributes, its derived operand buffers must be
bit-identical to a fresh preparation of that bank, and a captured step that owns the bank as a ring must reproduce
the step that rewrites five tensors."""
import os

import numpy as np
import pytest
import torch

from neighborretr_b200 import ops, selfcheck, synth
from neighborretr_b200.bank import BankRing
from neighborretr_b200.graph import FIELDS, GraphedHeadStep

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("x3", [False, True])
def test_ring_equals_the_reference_fifo_and_a_fresh_preparation(x3):
    M, nt, nv, d = 10, 8, 4, 64
    h0 = synth.make_batch(M, nt, nv, d=d, seed=1).to("cuda")
    ref = {"ind": h0.idx, "t": h0.text_feat, "v": h0.video_feat, "mt": h0.text_mask, "mv": h0.video_mask}
    ring = BankRing(ref["ind"], ref["t"], ref["v"], ref["mt"], ref["mv"], x3=x3, mlp_bf16=True, batch_rows=12)
    for step, b in enumerate((3, 3, 4, 10, 12, 1, 7)):            # wraps, B not dividing M, B = M, B > M
        hb = synth.make_batch(b, nt, nv, d=d, seed=10 + step).to("cuda")
        idx = hb.idx + 100 * (step + 1)
        ring.insert(idx, hb.text_feat, hb.video_feat, hb.text_mask, hb.video_mask)
        ref = {"ind": torch.cat([idx, ref["ind"]])[:M], "t": torch.cat([hb.text_feat, ref["t"]])[:M],
               "v": torch.cat([hb.video_feat, ref["v"]])[:M], "mt": torch.cat([hb.text_mask, ref["mt"]])[:M],
               "mv": torch.cat([hb.video_mask, ref["mv"]])[:M]}
        ex = ring.export()
        assert int(ring.head_dev.item()) == ring.head
        for name, key in (("mb_ind", "ind"), ("mb_feat_t", "t"), ("mb_feat_v", "v"), ("mb_mask_t", "mt"), ("mb_mask_v", "mv")):
            assert torch.equal(ex[name], ref[key]), (step, name)
        # derived buffers, rotated into reference order, against a fresh preparation of the reference bank
        for feat, mask, xn, xnT, mlp, n, role in ((ref["t"], ref["mt"], ring.xn_t, ring.xnT_t, ring.mlp_t, nt, ring.roles[0]),
                                                  (ref["v"], ref["mv"], ring.xn_v, ring.xnT_v, ring.mlp_v, nv, ring.roles[1])):
            P = ops.Prepared(feat, bf16=True, mask=mask, split=role)
            PT, _ = P.bwd_source(ops.NR_PREC_BF16X3 if x3 else ops.NR_PREC_BF16)
            assert torch.equal(torch.roll(xn, -ring.head, 0), P.xn_bf16), (step, "operand copy")
            got_T = torch.roll(xnT[:, :M * n].reshape(-1, M, n), -ring.head, 1).reshape(-1, M * n)
            assert torch.equal(got_T, PT[:, :M * n]), (step, "transposed copy")
            raw = mlp[ring.batch_rows * n:].reshape(M, n, d)
            assert torch.equal(torch.roll(raw, -ring.head, 0), feat.to(torch.bfloat16)), (step, "MLP operand")


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_captured_step_with_ring_matches_the_five_tensor_step(precision, monkeypatch):
    nt, nv, M, b = 24, 12, 96, 32
    bank = synth.make_bank(M, nt, nv)
    dev = torch.device("cuda", 0)
    outs = {}
    for use_ring in ("1", "0"):
        monkeypatch.setenv("NR_BANK_RING", use_ring)
        model = selfcheck.make_model(synth.default_config(), dev, precision)
        selfcheck.set_bank(model, bank, dev)
        h = synth.make_batch(b, nt, nv, seed=5).to(dev)
        g = GraphedHeadStep(model, [getattr(h, f) for f in FIELDS])
        assert (g.ring is not None) == (use_ring == "1")
        res = []
        for i in range(4):                                       # 4 x 32 new rows: the ring wraps past M = 96
            hb = synth.make_batch(b, nt, nv, seed=20 + i).to(dev)
            losses = g(*[getattr(hb, f) for f in FIELDS]).clone()
            res.append((losses, g.grads["text_feat"].clone(), model.text_weight_fc[0].weight.grad.clone()))
        outs[use_ring] = (res, {n: getattr(model, n).clone() for n in g.bank_names})
    for (l1, t1, w1), (l0, t0, w0) in zip(outs["1"][0], outs["0"][0]):
        np.testing.assert_allclose(l1.cpu().numpy(), l0.cpu().numpy(), rtol=2e-5)
        assert float((t1 - t0).norm() / t0.norm()) < 2e-3         # atomics / split-K order only
        assert float((w1 - w0).norm() / w0.norm()) < 2e-3
    for n in outs["1"][1]:
        assert torch.equal(outs["1"][1][n], outs["0"][1][n].to(outs["1"][1][n].dtype)), n


def test_assignment_hands_the_bank_back_to_the_attributes():
    nt, nv, M, b = 24, 12, 64, 32
    dev = torch.device("cuda", 0)
    model = selfcheck.make_model(synth.default_config(), dev, "bf16")
    selfcheck.set_bank(model, synth.make_bank(M, nt, nv), dev)
    h = synth.make_batch(b, nt, nv, seed=5).to(dev)
    g = GraphedHeadStep(model, [getattr(h, f) for f in FIELDS])
    g(*[getattr(h, f) for f in FIELDS])
    assert model.__dict__["_nr_ring_live"] and torch.equal(model.mb_ind[:b].cpu(), h.idx.cpu())
    new_ind = torch.arange(M, device=dev) + 5000
    model.mb_ind = new_ind                                        # what MemoryBankManager does (with all five)
    assert not model.__dict__["_nr_ring_live"]
    assert torch.equal(model.mb_ind, new_ind) and torch.equal(model.mb_feat_t[:b], h.text_feat)   # the others: ring rows
