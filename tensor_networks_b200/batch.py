"""`TensorTrainBatch`: many tensor trains of one shape, resident in HBM, core-major.

The reference has no batch API -- batches arise from its callers looping over
`inner` / `tt_svd_round` (GMRES Gram-Schmidt pytens/algs.py:2752-2757, structure
search pytens/search/partition.py:139-141, cross convergence pytens/cross/cross.py:403-404).
Core k of the whole batch is ONE CUDA tensor of shape (B, r_{k-1}, n_k, r_k), so item i
is slab i of every core and a multi-GPU shard is a slice of dimension 0.
"""

from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import TTBatchDescriptor, check
from .tt import TensorTrain, _require_cuda, _stream_ptr, workspace


class TensorTrainBatch:
    def __init__(self, cores: Sequence[torch.Tensor]):
        _require_cuda()
        cores = [c if c.is_contiguous() else c.contiguous() for c in cores]
        if not cores:
            raise ValueError("a TensorTrainBatch needs at least one core")
        B = cores[0].shape[0]
        for k, c in enumerate(cores):
            if c.dtype != torch.float64 or not c.is_cuda or c.dim() != 4 or c.shape[0] != B:
                raise ValueError(f"core {k}: need a (B, r, n, r') CUDA float64 tensor, got {tuple(c.shape)}")
        if cores[0].shape[1] != 1 or cores[-1].shape[3] != 1:
            raise AssertionError("boundary bond ranks must be 1")
        for k in range(len(cores) - 1):
            if cores[k].shape[3] != cores[k + 1].shape[1]:
                raise AssertionError(f"bond {k} does not chain")
        self.cores: List[torch.Tensor] = cores
        self.item_ranks: Optional[torch.Tensor] = None  # (B, d+1) after a rounding

    # ------------------------------------------------------------------ construction
    @classmethod
    def rand(cls, batch: int, shape: Sequence[int], ranks: Sequence[int], seed: Optional[int] = None,
             scaled: bool = True, device="cuda") -> "TensorTrainBatch":
        _require_cuda()
        d = len(shape)
        assert len(ranks) + 1 == d
        r = [1] + [int(x) for x in ranks] + [1]
        gen = torch.Generator(device=device)
        if seed is not None:
            gen.manual_seed(int(seed))
        cores = []
        for k in range(d):
            c = torch.randn((batch, r[k], int(shape[k]), r[k + 1]), dtype=torch.float64, device=device, generator=gen)
            if scaled:
                c *= 1.0 / math.sqrt(shape[k] * r[k + 1])
            cores.append(c)
        return cls(cores)

    @classmethod
    def from_items(cls, items: Sequence[TensorTrain]) -> "TensorTrainBatch":
        d = items[0].d
        return cls([torch.stack([it.cores[k] for it in items], dim=0) for k in range(d)])

    @classmethod
    def from_numpy(cls, items: Sequence[Sequence[np.ndarray]], device="cuda") -> "TensorTrainBatch":
        """items[i][k] = core k of item i as a (r, n, r') array."""
        d = len(items[0])
        return cls([torch.from_numpy(np.stack([np.asarray(it[k], dtype=np.float64) for it in items])).to(device)
                    for k in range(d)])

    def item(self, i: int) -> TensorTrain:
        """Item i as a TensorTrain (views; honours per-item ranks after a rounding)."""
        if self.item_ranks is None:
            return TensorTrain([c[i] for c in self.cores])
        rk = [int(x) for x in self.item_ranks[i].tolist()]
        out = []
        for k, c in enumerate(self.cores):
            rl, n, rr = rk[k], int(c.shape[2]), rk[k + 1]
            out.append(c[i].reshape(-1)[: rl * n * rr].view(rl, n, rr))
        return TensorTrain(out)

    def clone(self) -> "TensorTrainBatch":
        out = TensorTrainBatch([c.clone() for c in self.cores])
        out.item_ranks = None if self.item_ranks is None else self.item_ranks.clone()
        return out

    # ------------------------------------------------------------------ accessors
    @property
    def batch(self) -> int:
        return int(self.cores[0].shape[0])

    @property
    def d(self) -> int:
        return len(self.cores)

    @property
    def device(self) -> torch.device:
        return self.cores[0].device

    def shape(self) -> List[int]:
        return [int(c.shape[2]) for c in self.cores]

    def ranks(self) -> List[int]:
        """Capacity bond ranks (the shape of the storage)."""
        return [int(c.shape[3]) for c in self.cores[:-1]]

    def bond_ranks(self) -> List[int]:
        return [1] + self.ranks() + [1]

    def nbytes(self) -> int:
        return sum(c.numel() * 8 for c in self.cores)

    def descriptor(self) -> TTBatchDescriptor:
        return TTBatchDescriptor(self.batch, self.shape(), self.bond_ranks(), [c.data_ptr() for c in self.cores])

    def shard(self, rank: int, world: int) -> "TensorTrainBatch":
        """Contiguous block partition of the batch (views, no copy)."""
        from .sharding import shard_range

        lo, hi = shard_range(self.batch, rank, world)
        return TensorTrainBatch([c[lo:hi] for c in self.cores])

    def __add__(self, other: "TensorTrainBatch") -> "TensorTrainBatch":
        """Item-wise formal sum by block-diagonal rank growth (pytens/algs.py:1339-1353)."""
        if self.shape() != other.shape() or self.batch != other.batch:
            raise AssertionError("batched TT sum needs equal batch and mode sizes")
        d = self.d
        out = []
        for k, (a, b) in enumerate(zip(self.cores, other.cores)):
            if d == 1:
                out.append(a + b)
            elif k == 0:
                out.append(torch.cat([a, b], dim=3))
            elif k == d - 1:
                out.append(torch.cat([a, b], dim=1))
            else:
                c = torch.zeros((a.shape[0], a.shape[1] + b.shape[1], a.shape[2], a.shape[3] + b.shape[3]),
                                dtype=torch.float64, device=a.device)
                c[:, : a.shape[1], :, : a.shape[3]] = a
                c[:, a.shape[1]:, :, a.shape[3]:] = b
                out.append(c)
        return TensorTrainBatch(out)

    # ------------------------------------------------------------------ hot path
    def inner(self, other: "TensorTrainBatch", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B,) CUDA tensor of <self_i, other_i> -- TensorNetwork.inner per item
        (pytens/algs.py:585-587); one fused kernel when all bond ranks are <= 32."""
        if self.item_ranks is not None or other.item_ranks is not None:
            raise RuntimeError("inner on a rounded batch: call .item(i) (ranks differ per item)")
        if self.shape() != other.shape() or self.batch != other.batch:
            raise AssertionError("inner: batches differ in size or free indices")
        L = _lib.lib()
        da, db = self.descriptor(), other.descriptor()
        ws = workspace(L.ttb_inner_batched_workspace_bytes(da.ref(), db.ref()), self.device)
        if out is None:
            out = torch.empty(self.batch, dtype=torch.float64, device=self.device)
        check(L.ttb_inner_batched_f64(da.ref(), db.ref(), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
        return out

    def inner_scatter(self, other: "TensorTrainBatch", peer_ptrs: Sequence[int], offset: int) -> None:
        """<self_i, other_i> of this shard stored at peer_ptrs[r] + 8 (offset + i) for EVERY r: the peers' full-batch
        result arrays are mapped into this process (`sharding.PeerGather`), so the kernel's epilogue is the all-gather
        (`ttb_inner_batched_scatter_f64`).  The caller synchronises the ranks (PeerGather.barrier) before reading."""
        if self.item_ranks is not None or other.item_ranks is not None:
            raise RuntimeError("inner on a rounded batch: call .item(i) (ranks differ per item)")
        if self.shape() != other.shape() or self.batch != other.batch:
            raise AssertionError("inner: batches differ in size or free indices")
        import ctypes

        L = _lib.lib()
        da, db = self.descriptor(), other.descriptor()
        ws = workspace(L.ttb_inner_batched_scatter_workspace_bytes(da.ref(), db.ref()), self.device)
        ptrs = (ctypes.c_void_p * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
        check(L.ttb_inner_batched_scatter_f64(da.ref(), db.ref(), ptrs, len(peer_ptrs), int(offset), ws.data_ptr(), ws.numel(),
                                              _stream_ptr()))

    def norm(self) -> torch.Tensor:
        """(B,) CUDA tensor of sqrt(|<X_i, X_i>|) -- TensorNetwork.norm, pytens/algs.py:589-594."""
        return self.inner(self).abs().sqrt()

    def round(self, eps: float, max_rank: Optional[int] = None) -> "TensorTrainBatch":
        """Round every item in place with relative accuracy eps -- tt_svd_round per item
        (pytens/algs.py:1841-1903).  No host synchronisation: the per-item bond ranks land in
        `self.item_ranks` ((B, d+1) int64 CUDA tensor) and `self.round_status` ((B,) int32)."""
        if self.item_ranks is not None:
            raise RuntimeError("batch already rounded; rebuild it before rounding again")
        L = _lib.lib()
        desc = self.descriptor()
        ws = workspace(L.ttb_round_batched_workspace_bytes(desc.ref()), self.device)
        ranks = torch.empty((self.batch, self.d + 1), dtype=torch.int64, device=self.device)
        status = torch.zeros(self.batch, dtype=torch.int32, device=self.device)
        check(
            L.ttb_round_batched_f64(
                desc.ref(), float(eps), int(max_rank) if max_rank else 0, ranks.data_ptr(), status.data_ptr(),
                ws.data_ptr(), ws.numel(), _stream_ptr(),
            )
        )
        self.item_ranks = ranks
        self.round_status = status
        return self
