"""Batch sharding across GPUs: one process per GPU, `torch.distributed` for plumbing.

Only the batched workloads shard (independent tensor trains); a single large TT stays
on one GPU because its sweep is a strict recurrence over cores.  The batch is split into
contiguous blocks, every rank works on its block with no data-path collective, and the
per-item results (fp64 scalars, int64 rank tables) are all-gathered -- NCCL over
NVLink on GPUs, gloo on CPU for the tests of this host-side logic.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: the first (batch % world) ranks get one extra item."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, rem = divmod(int(batch), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def all_gather_items(local: torch.Tensor, batch: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Gather per-item results of every shard into the full (batch, ...) tensor on every rank.

    `local` holds this rank's block (shard_range order) along dimension 0.  Shards may
    differ by one item; they are padded to the largest shard for a single all_gather.
    """
    if not dist.is_initialized():
        if local.shape[0] != batch:
            raise ValueError("single process: local must hold the whole batch")
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(batch, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: expected {hi - lo} local items, got {local.shape[0]}")
    cap = -(-batch // world)
    tail = tuple(local.shape[1:])
    padded = torch.zeros((cap,) + tail, dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    gathered = torch.empty((world * cap,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    out = torch.empty((batch,) + tail, dtype=local.dtype, device=local.device)
    for r in range(world):
        l, h = shard_range(batch, r, world)
        out[l:h] = gathered[r * cap : r * cap + (h - l)]
    return out


def inner_sharded(a_local, b_local, batch: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """<A_i, B_i> for the whole batch: local fused kernel + all-gather of the scalars."""
    return all_gather_items(a_local.inner(b_local), batch, group)


def round_sharded(y_local, eps: float, batch: int, max_rank: Optional[int] = None,
                  group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Round every local item in place and all-gather the (batch, d+1) rank table.

    The rounded cores stay sharded (they are what the next local step consumes)."""
    y_local.round(eps, max_rank=max_rank)
    return all_gather_items(y_local.item_ranks, batch, group)
