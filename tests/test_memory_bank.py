"""Memory-bank prefill / reset (reference utils/memory_bank.py:22-268) — host data movement, CPU tests:
golden vectors from the reference's own MemoryBankManager (tests/golden/prefill.npz, oracle/gen_golden.py),
rank-major row order under gloo (world 2), ragged last batch, error behaviour."""
import logging
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import load_golden
from neighborretr_b200 import synth
from neighborretr_b200.memory_bank import MemoryBankManager


def _args(**kw):
    a = dict(logger=logging.getLogger("test"), mb_batch=4, batch_size=5, distributed=False, world_size=1, local_rank=0)
    a.update(kw)
    return SimpleNamespace(**a)


def test_prefill_matches_reference_golden():
    gold = load_golden("prefill")
    model = synth.ToyEncoder(d=8)
    mgr = MemoryBankManager(_args())
    n = mgr.load_memory_bank(model, synth.make_prefill_loader(6, 5), "cpu", 0)
    assert n == int(gold["rows"]) == 20 and model.mb_batch == int(gold["mb_batch"])
    assert not model.training                                        # left in eval mode like the reference (:103)
    for name, key in (("mb_ind", "ind"), ("mb_feat_t", "feat_t"), ("mb_feat_v", "feat_v"), ("mb_mask_t", "mask_t"),
                      ("mb_mask_v", "mask_v")):
        got = getattr(model, name)
        assert got.dtype == torch.from_numpy(gold[key]).dtype and np.array_equal(got.numpy(), gold[key]), name
        assert got.is_contiguous() or got.numel() == 0
    out = mgr.clear_memory_bank(model)
    assert out is model
    assert [model.mb_ind.numel(), model.mb_feat_t.dim(), model.mb_mask_v.dim(), model.mb_batch] == \
        gold["cleared_shapes"].tolist()
    assert model.mb_ind.dtype == torch.long and model.mb_feat_t.shape == (0, 0, 0) and model.mb_mask_t.shape == (0, 0)


def test_prefill_short_loader_ragged_batch_and_wrapped_model():
    """Fewer loader batches than mb_batch, a smaller last batch, and a DDP-style ``.module`` wrapper."""
    batches = synth.make_prefill_loader(3, 5)
    batches[-1] = tuple(t[:2] for t in batches[-1])
    inner = synth.ToyEncoder(d=8)
    wrapped = SimpleNamespace(module=inner)
    n = MemoryBankManager(_args(mb_batch=10)).load_memory_bank(wrapped, batches, "cpu", 3)
    assert n == 12 and inner.mb_feat_t.shape == (12, 4, 8) and inner.mb_mask_v.shape == (12, 3)
    want_t = torch.cat([inner.get_text_video_feat(b[0], b[1], b[2], b[3])[0] for b in batches]).detach()
    assert torch.equal(inner.mb_feat_t, want_t)
    assert inner.mb_ind.tolist() == torch.cat([b[4] for b in batches]).view(-1).tolist()
    assert not inner.mb_feat_t.requires_grad
    # a later batch LARGER than the first one still lands in order
    big = synth.make_prefill_loader(2, 5)
    big[0] = tuple(t[:2] for t in big[0])
    m2 = synth.ToyEncoder(d=8)
    assert MemoryBankManager(_args(mb_batch=2)).load_memory_bank(m2, big, "cpu", 0) == 7
    assert m2.mb_ind.tolist() == torch.cat([b[4] for b in big]).view(-1).tolist()
    # nothing to process -> 0 and the bank is left untouched
    m3 = synth.ToyEncoder(d=8)
    assert MemoryBankManager(_args(mb_batch=0)).load_memory_bank(m3, big, "cpu", 0) == 0 and m3.mb_batch == 0


def test_prefill_rejects_inconsistent_batches_and_missing_loader(monkeypatch):
    import sys
    monkeypatch.setitem(sys.modules, "NeighborRetr.dataloaders.data_dataloaders", None)   # reference not importable
    bad = synth.make_prefill_loader(2, 5)
    bad[1] = (bad[1][0][:, :3],) + bad[1][1:]                       # text ids with fewer words than the mask
    model = synth.ToyEncoder(d=8)
    with pytest.raises(ValueError, match="does not match the first batch"):
        MemoryBankManager(_args()).load_memory_bank(model, bad, "cpu", 0)
    with pytest.raises(RuntimeError, match="needs the reference's dataloaders"):
        MemoryBankManager(_args(datatype="msrvtt")).load_memory_bank(model, None, "cpu", 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = synth.ToyEncoder(d=8)
        args = _args(distributed=True, world_size=world, local_rank=rank, mb_batch=3)
        n = MemoryBankManager(args).load_memory_bank(model, synth.make_prefill_loader(4, 5, rank=rank), "cpu", 0)
        assert n == world * 15
        # reference :183-190: cat over this rank's batches, then rank-ordered gather
        want_ind, want_t, want_mv = [], [], []
        for r in range(world):
            bs = synth.make_prefill_loader(4, 5, rank=r)[:3]
            want_ind.append(torch.cat([b[4] for b in bs]).view(-1))
            want_t.append(torch.cat([model.get_text_video_feat(b[0], b[1], b[2], b[3])[0] for b in bs]).detach())
            want_mv.append(torch.cat([b[3] for b in bs]))
        assert torch.equal(model.mb_ind, torch.cat(want_ind))
        assert torch.equal(model.mb_feat_t, torch.cat(want_t))
        assert torch.equal(model.mb_mask_v, torch.cat(want_mv)) and model.mb_mask_v.dtype == torch.int64
        assert model.mb_feat_v.shape == (world * 15, 3, 8) and model.mb_batch == world * 15
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_prefill_world2_gloo_is_rank_major():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


@pytest.mark.gpu
def test_gpu_prefill_feeds_the_fifo():
    """Prefill on the device through the head's own ``get_text_video_feat`` hook, then the kernel FIFO
    (nr_fifo_update) keeps cat(new, old)[:capacity] on the prefilled buffers (reference modeling.py:222-249)."""
    from neighborretr_b200.modeling import NeighborRetr
    toy = synth.ToyEncoder(d=8)
    head = NeighborRetr(synth.default_config(), encoder=lambda *a: toy.get_text_video_feat(*a), width=8).cuda()
    toy.cuda()
    loader = synth.make_prefill_loader(6, 5)
    n = MemoryBankManager(_args()).load_memory_bank(head, loader, "cuda", 0)
    gold = load_golden("prefill")
    assert n == 20 and head.mb_feat_t.is_cuda and head.mb_feat_t.dtype == torch.float32
    assert np.array_equal(head.mb_ind.cpu().numpy(), gold["ind"])
    np.testing.assert_allclose(head.mb_feat_t.cpu().numpy(), gold["feat_t"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(head.mb_feat_v.cpu().numpy(), gold["feat_v"], rtol=1e-5, atol=1e-5)
    assert np.array_equal(head.mb_mask_t.cpu().numpy(), gold["mask_t"])
    old = {k: getattr(head, k).clone() for k in ("mb_ind", "mb_feat_t", "mb_feat_v", "mb_mask_t", "mb_mask_v")}
    b = tuple(t.cuda() for t in loader[5])
    ft, fv = head.get_text_video_feat(b[0], b[1], b[2], b[3])
    head.update_memory_bank(b[4].view(-1), ft, fv, b[1], b[3])
    assert torch.equal(head.mb_ind, torch.cat([b[4].view(-1), old["mb_ind"]])[:20])
    assert torch.equal(head.mb_feat_t, torch.cat([ft, old["mb_feat_t"]])[:20])
    assert torch.equal(head.mb_feat_v, torch.cat([fv, old["mb_feat_v"]])[:20])
    assert torch.equal(head.mb_mask_t, torch.cat([b[1], old["mb_mask_t"]])[:20])
    assert torch.equal(head.mb_mask_v, torch.cat([b[3], old["mb_mask_v"]])[:20])
    MemoryBankManager(_args()).clear_memory_bank(head)
    assert head.mb_batch == 0 and head.mb_feat_v.shape == (0, 0, 0) and head.mb_feat_v.is_cuda
