"""Multi-GPU plumbing for the rollout: trajectories are independent, so the batch is split into
contiguous shards (one process per GPU, `torch.distributed`), maps are replicated and the ONLY
collectives are:

  * planning / shooting: one all-gather of the per-trajectory costs (B floats per rank), then a local argmin
    (`monoforce_ros/nodes/monoforce_node.py:91,126`);
  * training on a shared map: all-reduce of the two map gradients.

The functions work with any initialised backend (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_trajectories: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous shard `rank` owns; remainders go to the lowest ranks."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, rem = divmod(n_trajectories, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


def gather_costs(local_costs: torch.Tensor, n_trajectories: int) -> torch.Tensor:
    """All ranks get the (n_trajectories,) cost vector in global trajectory order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_costs
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n_trajectories, r, world) for r in range(world)]
    if all(hi - lo == sizes[0][1] - sizes[0][0] for lo, hi in sizes):
        out = torch.empty(n_trajectories, dtype=local_costs.dtype, device=local_costs.device)
        dist.all_gather_into_tensor(out, local_costs.contiguous())
        return out
    # uneven shards: pad every contribution to the largest shard, gather, drop the padding
    width = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros(width, dtype=local_costs.dtype, device=local_costs.device)
    padded[: local_costs.shape[0]] = local_costs
    out = torch.empty(world * width, dtype=local_costs.dtype, device=local_costs.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def best_trajectory(local_costs: torch.Tensor, n_trajectories: int) -> Tuple[int, float]:
    """Global index and cost of the cheapest trajectory (what the planner publishes)."""
    costs = gather_costs(local_costs, n_trajectories)
    i = int(torch.argmin(costs))
    return i, float(costs[i])


def _adjacent_view(tensors):
    """One flat view over `tensors` when they are contiguous, back to back and in one storage; else None."""
    first = tensors[0]
    if not all(t.is_contiguous() for t in tensors):
        return None
    store = first.untyped_storage()
    off = first.storage_offset()
    n = 0
    for t in tensors:
        if t.untyped_storage().data_ptr() != store.data_ptr() or t.storage_offset() != off + n:
            return None
        n += t.numel()
    return torch.as_strided(first, (n,), (1,), off)


def allreduce_map_grads(*grads: torch.Tensor) -> None:
    """Sum the shared-map gradients over ranks (in place).  Several maps of one dtype travel as ONE flat all-reduce
    (the collective is latency-bound at 256 KiB per map: one launch instead of one per map)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    live = [g for g in grads if g is not None]
    if len(live) <= 1 or len({(g.dtype, g.device) for g in live}) != 1:
        for g in live:
            dist.all_reduce(g)
        return
    flat = _adjacent_view(live)
    if flat is not None:          # DPhysics' backward hands out both map gradients as halves of one buffer: no copies at all
        dist.all_reduce(flat)
        return
    flat = torch.cat([g.reshape(-1) for g in live])
    dist.all_reduce(flat)
    off = 0
    for g in live:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
