"""Batch sharding over the GPUs of one box: one process per GPU, no collective on the data path.

Units (slot x UE x rx antenna) are independent (the reference call owns all of its accumulators,
/root/reference src/ce_rule_tensorized.py:790-796), so rank r simply takes a contiguous block of
the flattened unit index, aligned to ``pilot_share`` so that the rx antennas of one UE stay on one
GPU.  The only exchange is the final gather of the per-unit scalar table (48 B per unit); the
channel grids stay sharded on the GPU that produced them.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """[begin, end) of rank's block: blocks are contiguous, multiples of ``align`` and differ by at
    most one aligned group."""
    if n_units % align:
        raise ValueError("n_units must be a multiple of align")
    groups = n_units // align
    base, extra = divmod(groups, world)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return begin * align, end * align


def gather_scalars(local: torch.Tensor, n_units: int, align: int = 1,
                   group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the [local_units, 6] float64 scalar tables into the global [n_units, 6] table
    (NCCL over NVLink for CUDA tensors, gloo for CPU tensors in the tests).  Shards may differ by
    one aligned group, so they are padded to the largest shard for the collective."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_units, r, world, align) for r in range(world)]
    longest = max(e - b for b, e in sizes)
    padded = torch.zeros((longest, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    table = torch.empty((world * longest, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(table, padded, group=group)
    return torch.cat([table[r * longest: r * longest + (e - b)] for r, (b, e) in enumerate(sizes)], dim=0)


def estimate_sharded(plan, rx_local: torch.Tensor, pilots_local: torch.Tensor, beta: float, pilot_share: int,
                     n_units_global: int, group: Optional[dist.ProcessGroup] = None, out=None):
    """Run this rank's block through the fused kernel and gather the scalar table.
    Returns (local channel grids, global scalar table)."""
    grid, sc = plan.estimate(rx_local, pilots_local, beta, pilot_share, out=out)
    return grid, gather_scalars(sc, n_units_global, pilot_share, group)
