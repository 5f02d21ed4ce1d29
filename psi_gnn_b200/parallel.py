"""Graph-sharded multi-GPU execution: one process per GPU, each rank solves its own sub-batch, gradients are averaged.

This is the reference's only parallelism (PyG ``DataParallel`` at dirichlet/psignn/main.py:106: the data list is split by
cumulative node count, each replica runs its *own* forward and backward Broyden solve on its sub-batch, gradients are
summed on device 0) re-expressed as SPMD ranks: no communication inside the solver, one ``all_reduce`` of the flat
gradient (1 444 floats Dirichlet / 2 175 mixed) per step.  Works with ``nccl`` (GPU) and ``gloo`` (CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist

from .synthetic import GraphData, split_graphs


def shard_batch(batch: GraphData, rank: int, world_size: int) -> GraphData:
    """the rank's contiguous group of graphs, balanced by node count (PyG DataParallel.scatter semantics)"""
    if world_size == 1:
        return batch
    if batch.num_graphs < world_size:
        raise ValueError("cannot shard %d graphs over %d ranks" % (batch.num_graphs, world_size))
    return split_graphs(batch, world_size)[rank]


def flatten_grads(params: Sequence[torch.nn.Parameter]) -> torch.Tensor:
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])


def unflatten_grads_(params: Sequence[torch.nn.Parameter], flat: torch.Tensor) -> None:
    o = 0
    for p in params:
        n = p.numel()
        g = flat[o:o + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        o += n


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: int = None, group=None) -> None:
    """average the gradients over the ranks with ONE collective on the flat blob (losses are means over each shard and the
    reference averages the replica losses, training_class.py:156-159)."""
    if not dist.is_available() or not dist.is_initialized():
        return
    ws = world_size or dist.get_world_size(group)
    if ws == 1:
        return
    params = [p for p in params if p.requires_grad]
    flat = flatten_grads(params)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(ws)
    unflatten_grads_(params, flat)


def allreduce_scalars(values: List[float], device, op=None, group=None) -> List[float]:
    """reduce a handful of logging scalars (loss terms, timings) in one call"""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=group)
    return t.tolist()
