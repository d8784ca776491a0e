"""Batch sharding across GPUs: one process per GPU, `torch.distributed` for plumbing.

Only the batched workloads shard (independent tensor trains); a single large TT stays
on one GPU because its sweep is a strict recurrence over cores.  The batch is split into
contiguous blocks, every rank works on its block with no data-path collective, and the
per-item results -- fp64 scalars, int64 rank tables and, on request, the rounded cores --
are all-gathered: NCCL over NVLink on GPUs, gloo on CPU for the tests of this host-side logic.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

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


def all_gather_items(local: torch.Tensor, batch: int, group: Optional[dist.ProcessGroup] = None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Gather per-item results of every shard into the full (batch, ...) tensor on every rank.

    `local` holds this rank's block (shard_range order) along dimension 0.  When the batch divides
    evenly (the usual case) this is ONE collective straight into the result -- no staging copies.
    Uneven shards are padded to the largest one; the blocks of the longer shards then already sit in
    place in the gathered array and only the shorter ones are moved.
    """
    if not dist.is_initialized():
        if local.shape[0] != batch:
            raise ValueError("single process: local must hold the whole batch")
        if out is not None and out.data_ptr() != local.data_ptr():
            out.copy_(local)  # same contract as the collective path: the caller's buffer holds the result
            return out
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(batch, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: expected {hi - lo} local items, got {local.shape[0]}")
    tail = tuple(local.shape[1:])
    local = local.contiguous()
    if batch % world == 0:
        if out is None:
            out = torch.empty((batch,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    cap = -(-batch // world)
    padded = torch.zeros((cap,) + tail, dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    gathered = torch.empty((world * cap,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    rem = batch % world  # ranks < rem hold cap items: their blocks are already contiguous at the front
    if out is None:
        out = torch.empty((batch,) + tail, dtype=local.dtype, device=local.device)
    out[: rem * cap] = gathered[: rem * cap]
    for r in range(rem, world):
        l, h = shard_range(batch, r, world)
        out[l:h] = gathered[r * cap : r * cap + (h - l)]
    return out


class PeerGather:
    """Full-batch result array that the kernels of EVERY rank write into directly over NVLink.

    The array lives in symmetric memory (`torch.distributed._symmetric_memory`: one allocation per rank, each mapped
    into all processes of the node), so a sharded kernel stores result i of its shard at `ptrs[r] + offset + i` for
    every rank r as its own epilogue -- compute and all-gather are ONE kernel and the only cross-rank step left is a
    signal-pad barrier (`barrier()`, a few microseconds) instead of an NCCL collective behind the kernel (measured
    at N = 8 on configs[4]: ~0.35 ms per step for 8 KiB of payload).  `ptrs` is None when symmetric memory is not
    available (CPU / gloo, peer access refused, a single process without CUDA): callers then use the NCCL path.
    A buffer must not be written again before every rank has finished reading it (second `barrier()`, or two buffers).
    """

    def __init__(self, batch: int, dtype: torch.dtype = torch.float64, device=None, group: Optional[dist.ProcessGroup] = None):
        self.batch = int(batch)
        self.group = group
        self.handle = None
        self.ptrs: Optional[List[int]] = None
        self.why_not: Optional[str] = None
        multi = dist.is_initialized() and dist.get_world_size(group) > 1
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        device = torch.device(device)
        if not multi:
            self.tensor = torch.empty(self.batch, dtype=dtype, device=device)
            if device.type == "cuda":
                self.ptrs = [self.tensor.data_ptr()]
            return
        if device.type != "cuda" or dist.get_world_size(group) > 8:
            self.tensor = torch.empty(self.batch, dtype=dtype, device=device)
            self.why_not = "symmetric memory needs CUDA devices of one node (at most 8 ranks)"
            return
        # Two steps, with agreement in between: the allocation is local and may fail on one rank only (then NO rank
        # may enter the rendezvous, which is a collective over the group's store and would wait for the missing rank).
        buf, err = None, None
        try:
            import torch.distributed._symmetric_memory as symm_mem

            buf = symm_mem.empty(self.batch, dtype=dtype, device=device)
        except Exception as exc:
            err = f"{type(exc).__name__}: {exc}"
        ok = torch.tensor([1 if buf is not None else 0], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            try:
                self.handle = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
                self.ptrs = [int(x) for x in self.handle.buffer_ptrs]
                self.tensor = buf
            except Exception as exc:  # no peer mapping on this box (the same on every rank): the collective path
                err = f"{type(exc).__name__}: {exc}"
                self.handle, self.ptrs = None, None
        if self.ptrs is None:
            self.tensor = torch.empty(self.batch, dtype=dtype, device=device)
            self.why_not = err or "symmetric memory allocation failed on another rank"

    @property
    def fused(self) -> bool:
        """True when kernels can store into every rank's array (or there is only one rank)."""
        return self.ptrs is not None

    def barrier(self) -> None:
        """Every rank's stores issued before this point (on the current stream) are visible to every rank after it."""
        if self.handle is not None:
            self.handle.barrier()


def inner_sharded(a_local, b_local, batch: int, group: Optional[dist.ProcessGroup] = None,
                  gather: Optional[PeerGather] = None) -> torch.Tensor:
    """<A_i, B_i> for the whole batch on every rank.  With a `PeerGather` whose buffers are peer-mapped the fused
    kernel stores its results into every rank's array itself (one kernel + a signal barrier); otherwise local
    kernel + NCCL all-gather of the scalars."""
    if gather is not None and gather.fused:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        lo, hi = shard_range(batch, rank, world)
        if a_local.batch != hi - lo:
            raise ValueError(f"rank {rank}: expected {hi - lo} local items, got {a_local.batch}")
        a_local.inner_scatter(b_local, gather.ptrs, lo)
        gather.barrier()
        return gather.tensor
    return all_gather_items(a_local.inner(b_local), batch, group, out=gather.tensor if gather is not None else None)


def round_sharded(y_local, eps: float, batch: int, max_rank: Optional[int] = None,
                  group: Optional[dist.ProcessGroup] = None, gather_cores: bool = False, arena=None):
    """Round every local item in place and all-gather the (batch, d+1) rank table.

    By default the rounded cores stay sharded (they are what the next local step consumes).  With
    `gather_cores` the cores are all-gathered as well and a full `TensorTrainBatch` is returned next to
    the table (north_star item 4: "all-gather of scalar and core results"); with an `arena` (see all_gather_cores) the
    pack kernel gathers them over NVLink peer memory itself."""
    y_local.round(eps, max_rank=max_rank)
    table = all_gather_items(y_local.item_ranks, batch, group)
    if not gather_cores:
        return table
    return table, all_gather_cores(y_local, batch, table, group, arena=arena)


def padded_ranks(table: torch.Tensor) -> List[int]:
    """Bond ranks of the uniform layout: the largest rank any item of the WHOLE batch has per bond."""
    return [int(x) for x in table.max(dim=0).values.tolist()]


def pack_rounded(y_local, rcap: Sequence[int]) -> List[torch.Tensor]:
    """Cores of a rounded local batch in the uniform zero-padded layout (nloc, rcap[k], n_k, rcap[k+1])
    (`ttb_pack_rounded_cores_f64`); zero padding keeps every item the same tensor."""
    from . import _lib
    from ._lib import check
    from .tt import _stream_ptr

    if y_local.item_ranks is None:
        raise RuntimeError("pack_rounded: the batch has not been rounded")
    L = _lib.lib()
    d, nloc = y_local.d, y_local.batch
    out = []
    for k, c in enumerate(y_local.cores):
        n = int(c.shape[2])
        slab = int(c.shape[1]) * n * int(c.shape[3])
        dst = torch.empty((nloc, int(rcap[k]), n, int(rcap[k + 1])), dtype=torch.float64, device=c.device)
        check(L.ttb_pack_rounded_cores_f64(c.data_ptr(), nloc, slab, n, y_local.item_ranks.data_ptr(), d, k,
                                           int(rcap[k]), int(rcap[k + 1]), dst.data_ptr(), _stream_ptr()))
        out.append(dst)
    return out


def all_gather_padded_cores(cores_local: Sequence[torch.Tensor], batch: int,
                            group: Optional[dist.ProcessGroup] = None) -> List[torch.Tensor]:
    """All-gather uniformly laid out cores (one collective per core, straight into the result)."""
    return [all_gather_items(c, batch, group) for c in cores_local]


def gathered_cores_numel(y_local, batch: int, table: torch.Tensor) -> int:
    """Elements of the uniform padded layout of the whole batch (the size of the `arena` of all_gather_cores)."""
    rcap = padded_ranks(table)
    return sum(int(batch) * rcap[k] * int(c.shape[2]) * rcap[k + 1] for k, c in enumerate(y_local.cores))


def all_gather_cores(y_local, batch: int, table: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                     arena: Optional[PeerGather] = None):
    """The rounded cores of every shard on every rank, as a `TensorTrainBatch` whose bond ranks are the
    batch-wide maxima (items with smaller ranks are zero-padded; cfg5: 8192 items, ranks 16 -> 2.3 GB instead of
    the 9.7 GB of raw slabs).

    With an `arena` (a `PeerGather` of at least gathered_cores_numel() doubles whose buffers are peer-mapped) the pack
    kernel of every core stores its output straight into EVERY rank's arena over NVLink
    (`ttb_pack_rounded_cores_scatter_f64`): pack and all-gather are one kernel per core and the ranks meet at one
    signal barrier at the end; the returned cores are views of the arena (valid until it is written again).
    Otherwise: pack kernel + one NCCL all-gather per core."""
    from .batch import TensorTrainBatch

    rcap = padded_ranks(table)
    if arena is not None and arena.fused:
        import ctypes

        from . import _lib
        from ._lib import check
        from .tt import _stream_ptr

        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        lo, hi = shard_range(batch, rank, world)
        if y_local.batch != hi - lo:
            raise ValueError(f"rank {rank}: expected {hi - lo} local items, got {y_local.batch}")
        if y_local.item_ranks is None:
            raise RuntimeError("all_gather_cores: the batch has not been rounded")
        if arena.tensor.numel() < gathered_cores_numel(y_local, batch, table) or arena.tensor.dtype != torch.float64:
            raise ValueError("all_gather_cores: arena too small (see gathered_cores_numel) or not float64")
        L = _lib.lib()
        d, nloc = y_local.d, y_local.batch
        full, off = [], 0
        for k, c in enumerate(y_local.cores):
            n = int(c.shape[2])
            slab = int(c.shape[1]) * n * int(c.shape[3])
            cnt = int(batch) * rcap[k] * n * rcap[k + 1]
            ptrs = (ctypes.c_void_p * len(arena.ptrs))(*[int(p) + 8 * off for p in arena.ptrs])
            check(L.ttb_pack_rounded_cores_scatter_f64(c.data_ptr(), nloc, slab, n, y_local.item_ranks.data_ptr(), d, k,
                                                       rcap[k], rcap[k + 1], ptrs, len(arena.ptrs), lo, _stream_ptr()))
            full.append(arena.tensor[off : off + cnt].view(int(batch), rcap[k], n, rcap[k + 1]))
            off += cnt
        arena.barrier()
        return TensorTrainBatch(full)
    full = all_gather_padded_cores(pack_rounded(y_local, rcap), batch, group)
    return TensorTrainBatch(full)
