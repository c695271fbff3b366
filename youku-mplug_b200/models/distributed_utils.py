"""Differentiable collectives used by the contrastive branches.

Mirrors `models/distributed_utils.py:all_gather` (:87-96 -> _AllGather :285-311): forward gathers one
tensor per rank, backward reduce-scatters (sum) the incoming gradients.  Implemented with ONE
all_gather_into_tensor into a contiguous [W*B, ...] buffer (the reference allocates W tensors and
concatenates) and one reduce_scatter_tensor in backward; works on NCCL
(GPU) and on gloo (CPU tests; reduce_scatter is emulated with all_reduce there).
"""
import torch
import torch.distributed as dist
from torch.autograd import Function
from torch.distributed import group


class _AllGatherContig(Function):
    @staticmethod
    def forward(ctx, grp, tensor):
        tensor = tensor.contiguous()
        ctx.group = grp
        world = dist.get_world_size(group=grp)
        out = torch.empty((world,) + tuple(tensor.shape), dtype=tensor.dtype, device=tensor.device)
        dist.all_gather_into_tensor(out.view(-1), tensor.view(-1), group=grp)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        grp = ctx.group
        grad_out = grad_out.contiguous()
        world = dist.get_world_size(group=grp)
        rank = dist.get_rank(group=grp)
        gx = torch.empty_like(grad_out[0])
        if dist.get_backend(group=grp) == dist.Backend.NCCL:
            dist.reduce_scatter_tensor(gx.view(-1), grad_out.view(-1), op=dist.ReduceOp.SUM, group=grp)
        else:
            g = grad_out.clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=grp)
            gx.copy_(g[rank])
        return None, gx


def all_gather_cat(tensor, grp=group.WORLD):
    """[B, ...] per rank -> [W*B, ...] (== torch.cat(all_gather(tensor), 0)), differentiable."""
    out = _AllGatherContig.apply(grp, tensor)
    return out.view((-1,) + tuple(tensor.shape[1:]))


def all_gather(tensor, group=group.WORLD):
    """Reference-compatible signature: returns a tuple of per-rank tensors (views of one buffer)."""
    out = _AllGatherContig.apply(group, tensor)
    return tuple(out.unbind(0))


@torch.no_grad()
def concat_all_gather(tensor):
    """Non-differentiable gather + concat (models/distributed_gpt3.py:1221-1231); bit-exact for ids."""
    world = dist.get_world_size()
    out = torch.empty((world,) + tuple(tensor.shape), dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(out.view(-1), tensor.contiguous().view(-1))
    return out.view((-1,) + tuple(tensor.shape[1:]))
