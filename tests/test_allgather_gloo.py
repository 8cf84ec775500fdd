"""world_size-2 gloo tests of the cross-rank gather on CPU (the N>1 host path; reference until_module.py:367-412):
rank-ordered concatenation, backward = local slice (AllGather) or summed slice (AllGather2), int64 payloads."""
import os
import socket
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neighborretr_b200.until_module import AllGather, AllGather2
    args = SimpleNamespace(world_size=world, local_rank=rank)
    try:
        b = 3
        x = (torch.arange(b * 4, dtype=torch.float32).reshape(b, 4) + 100 * rank).requires_grad_(True)
        y = AllGather.apply(x, args)
        assert y.shape == (world * b, 4)
        for r in range(world):
            assert torch.equal(y[r * b:(r + 1) * b], torch.arange(b * 4, dtype=torch.float32).reshape(b, 4) + 100 * r)
        w = torch.arange(world * b * 4, dtype=torch.float32).reshape(world * b, 4) * (rank + 1)
        (y * w).sum().backward()
        assert torch.equal(x.grad, w[rank * b:(rank + 1) * b])          # slice, no reduction
        idx = torch.arange(b, dtype=torch.int64) + 10 * rank
        gi = AllGather.apply(idx, args)
        assert gi.dtype == torch.int64 and gi.tolist() == [0, 1, 2, 10, 11, 12]
        x2 = x.detach().clone().requires_grad_(True)
        y2 = AllGather2.apply(x2, args)
        (y2 * w).sum().backward()
        tot = sum(torch.arange(world * b * 4, dtype=torch.float32).reshape(world * b, 4) * (r + 1) for r in range(world))
        assert torch.equal(x2.grad, tot[rank * b:(rank + 1) * b])        # summed over ranks, then sliced
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_allgather_world2_gloo():
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


def _exchange_worker(rank, world, port, q):
    """The block bookkeeping of the sharded head's exchange design (sharded.py): column blocks of S assembled from
    the other ranks' row blocks by an all-to-all of [b,b] blocks, and the inverse exchange of their gradients."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neighborretr_b200.sharded import _all_to_all_blocks, _reduce_scatter
    try:
        b, B, lo = 3, 3 * world, 3 * rank
        g = torch.Generator().manual_seed(5)
        S = torch.randn(B, B, generator=g)                                   # the full similarity matrix
        P = S[lo:lo + b].contiguous()                                        # what this rank computes: its text rows
        PT = P.t().contiguous()                                              # the kernel's second output [B, b]
        recv = _all_to_all_blocks(PT.view(world, b, b))                      # [q, v_l, a]
        S_col = torch.empty(b, B)
        S_col.view(b, world, b).copy_(recv.permute(1, 0, 2))
        assert torch.equal(S_col, S[:, lo:lo + b].t())                       # rows of S^T owned by this rank
        # backward: dS_col[v_l, a] belongs to the text row a of its owner
        dfull_row = torch.randn(B, B, generator=g)                           # d/dS from the t2v direction, by row owner
        dfull_col = torch.randn(B, B, generator=g)                           # d/dS^T from the v2t direction, by row owner
        dS_row = dfull_row[lo:lo + b]
        dS_col = dfull_col[lo:lo + b]                                        # [v_l, a]: gradient w.r.t. S[a, lo + v_l]
        back = _all_to_all_blocks(dS_col.reshape(b, world, b).permute(1, 0, 2))          # [r, v, a]
        dP = dS_row.reshape(b, world, b) + back.permute(2, 0, 1)
        want = dfull_row[lo:lo + b] + dfull_col.t()[lo:lo + b]               # dS[a, j] + dS^T[j, a] for this rank's rows a
        assert torch.allclose(dP.reshape(b, B), want)
        # reduce-scatter of per-rank partial gradients keeps this rank's rows of the sum
        part = torch.arange(B * 2, dtype=torch.float32).reshape(B, 2) * (rank + 1)
        tot = sum(torch.arange(B * 2, dtype=torch.float32).reshape(B, 2) * (r + 1) for r in range(world))
        assert torch.equal(_reduce_scatter(part, b), tot[lo:lo + b])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_exchange_blocks_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
