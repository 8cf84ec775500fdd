"""Loss modules and the cross-rank gather of the retrieval head — same names and call signatures as the
reference's NeighborRetr/models/until_module.py:56-412, running on the libnrhead.so kernels.

    CentralityWeightingLoss()(similarity_matrix, centrality_weights)
    NeighborAdjustingLoss()(similarity_matrix, memory_bank_matrix, num_neighbors, temperature)
    UniformRegularizationLoss()(similarity_matrix, logit_scale, beta=0.3, num_iterations=50)
    KLDivergenceLoss()(global_similarity, local_similarity)
    AllGather.apply(tensor, args) / AllGather2.apply(tensor, args)

Each module is parameter-free, takes ``config=None`` and returns a 0-dim tensor with autograd, exactly like
the reference's.  Inputs must be CUDA tensors: there is no CPU path.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._lib import NR_LOSS_CENTRALITY, NR_LOSS_KL, NR_LOSS_NEIGHBOR, NR_LOSS_UNIFORM


def _square(m, what):
    if m.dim() != 2 or m.shape[0] != m.shape[1]:
        raise RuntimeError(f"{what}: expected a square [B,B] matrix, got {tuple(m.shape)}")
    return m.shape[0]


class CentralityWeightingLoss(nn.Module):
    """reference until_module.py:294-328: mean_a( -w[a] * log_softmax(X)[a,a] )."""

    def __init__(self, config=None):
        super().__init__()

    def forward(self, similarity_matrix, centrality_weights):
        b = _square(similarity_matrix, "CentralityWeightingLoss")
        sums, _ = ops.row_losses(similarity_matrix, w=centrality_weights, flags=NR_LOSS_CENTRALITY)
        return sums[0] / b


class NeighborAdjustingLoss(nn.Module):
    """reference until_module.py:56-211.  ``last_neighbors`` keeps the [B,k] top-k columns picked by the
    kernel (descending similarity, ties -> lower column)."""

    def __init__(self, config=None):
        super().__init__()
        self.last_neighbors = None

    def forward(self, similarity_matrix, memory_bank_matrix, num_neighbors, temperature):
        b = _square(similarity_matrix, "NeighborAdjustingLoss")
        if b < num_neighbors + 2:
            # the reference raises IndexError / produces 9e15 garbage here
            raise IndexError(f"NeighborAdjustingLoss needs batch >= num_neighbors + 2 (got {b}, k={num_neighbors})")
        cbank = ops.row_mean(memory_bank_matrix)
        sums, nbr = ops.row_losses(similarity_matrix, cbank=cbank, k=num_neighbors, tau_nbr=temperature,
                                   flags=NR_LOSS_NEIGHBOR)
        self.last_neighbors = nbr
        return sums[1] / b


class UniformRegularizationLoss(nn.Module):
    """reference until_module.py:214-291: no-grad log-Sinkhorn target, cross-entropy of log_softmax(G*scale)."""

    def __init__(self, config=None):
        super().__init__()

    def sinkhorn_algorithm(self, scores, beta=0.3, num_iterations=50):
        """Transport-plan targets beta*Q + (1-beta)*I as a dense matrix (reference :222-266)."""
        b = _square(scores, "sinkhorn_algorithm")
        g = scores.detach().float().contiguous()
        u, v, _, _ = ops.sinkhorn_duals(g, g.t().contiguous(), num_iterations)
        nu = -torch.log(torch.tensor(2.0 * b, device=g.device))
        q = torch.exp(g + u[:, None] + v[None, :] - nu)
        return beta * q + (1 - beta) * torch.eye(b, device=g.device)

    def forward(self, similarity_matrix, logit_scale, beta=0.3, num_iterations=50):
        # logit_scale is the reference's `temperature` hyper-parameter (a Python float at every call site,
        # modeling.py:440-442); a tensor is accepted but costs a device synchronisation here — the fused head
        # (fused.HeadFunction) takes it from the config and never syncs
        b = _square(similarity_matrix, "UniformRegularizationLoss")
        g = similarity_matrix
        u, v, _, _ = ops.sinkhorn_duals(g, g.detach().t().contiguous(), num_iterations)
        sums, _ = ops.row_losses(g, G=g, sk_u=u, sk_v=v, tau_uni=float(logit_scale), beta=beta,
                                 flags=NR_LOSS_UNIFORM)
        return sums[3] / b


class KLDivergenceLoss(nn.Module):
    """reference until_module.py:331-359: F.kl_div(log_softmax(G), softmax(S), 'mean') (target not detached)."""

    def __init__(self, config=None):
        super().__init__()

    def forward(self, global_similarity, local_similarity):
        b = _square(local_similarity, "KLDivergenceLoss")
        sums, _ = ops.row_losses(local_similarity, G=global_similarity, flags=NR_LOSS_KL)
        return sums[2] / (b * b)


# ------------------------------------------------------------------------------------------------
# cross-rank gather (reference until_module.py:367-412)
# ------------------------------------------------------------------------------------------------
def _rank_of(args):
    # the reference slices by args.local_rank (single-node assumption); use the global
    # rank when the process group is up — identical on one node.
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank()
    return getattr(args, "local_rank", 0)


def _gather_contiguous(tensor, world_size):
    """One all_gather_into_tensor on a preallocated [W*b, ...] buffer (no list + cat copy)."""
    tensor = tensor.contiguous()
    out = torch.empty((world_size * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype,
                      device=tensor.device)
    torch.distributed.all_gather_into_tensor(out, tensor)
    return out


class AllGather(torch.autograd.Function):
    """Forward: rank-ordered concatenation along dim 0.  Backward: this rank's rows of the gradient."""

    @staticmethod
    def forward(ctx, tensor, args):
        ctx.rank = _rank_of(args)
        ctx.batch_size = tensor.shape[0]
        if args.world_size == 1:
            return tensor
        return _gather_contiguous(tensor, args.world_size)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output[ctx.batch_size * ctx.rank: ctx.batch_size * (ctx.rank + 1)], None


class AllGather2(torch.autograd.Function):
    """As AllGather, but the backward sums the gathered gradient over ranks before slicing
    (reference :391-412) — done as a reduce_scatter so each rank only receives its own rows."""

    @staticmethod
    def forward(ctx, tensor, args):
        ctx.rank = _rank_of(args)
        ctx.batch_size = tensor.shape[0]
        ctx.world_size = args.world_size
        if args.world_size == 1:
            return tensor
        return _gather_contiguous(tensor, args.world_size)

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.world_size == 1:
            return grad_output, None
        grad_output = grad_output.contiguous()
        out = torch.empty((ctx.batch_size,) + tuple(grad_output.shape[1:]), dtype=grad_output.dtype,
                          device=grad_output.device)
        torch.distributed.reduce_scatter_tensor(out, grad_output, op=torch.distributed.ReduceOp.SUM)
        return out, None
