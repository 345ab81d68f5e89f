"""Data-parallel helpers: one process per GPU, image pairs sharded contiguously over ranks, weights
replicated, no data-path collective for inference.  torch.distributed is the plumbing."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r owns pairs [r*n/world, (r+1)*n/world) (balanced contiguous split, earlier ranks take the
    remainder)."""
    if world <= 0 or not (0 <= rank < world) or n_pairs < 0:
        raise ValueError(f"bad shard request n={n_pairs} rank={rank} world={world}")
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction for multi-GPU numbers: the slowest rank defines the step time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_pairs(local: torch.Tensor, n_pairs: int) -> torch.Tensor:
    """All-gather per-rank results (possibly ragged shards) back into batch order.  Only needed when the
    caller wants every flow on every rank; inference itself needs no collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_pairs, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)


def allreduce_gradients(grad_flat: torch.Tensor, scalars: torch.Tensor = None, group=None) -> int:
    """The one data-path collective of a training step: SUM all-reduce of the flat fp32 gradient
    (20.1 MB for the default net) over the ranks, and the MEAN of the logged scalars (loss, EPE).  Returns the
    world size; the caller folds 1/world into the optimizer's gradient scale, so every rank applies the update
    of the global batch's mean-over-ranks loss -- N ranks x B pairs behave like the reference at batch B with
    the gradient averaged over ranks.  No-op (returns 1) outside a process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world == 1:
        return 1
    dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM, group=group)
    if scalars is not None:
        dist.all_reduce(scalars, op=dist.ReduceOp.SUM, group=group)
        scalars /= world
    return world
