"""`TensorTrain`: a device-resident tensor train driven through the ttb200 C ABI.

The reference has no such class -- a TT there is a `TensorNetwork` whose nodes
are the ints 0..d-1 in a chain (pytens/algs.py:1180-1218).  This container holds
the same data, core k as a CUDA fp64 tensor of shape (r_{k-1}, n_k, r_k) in
C order (byte-identical to the reference's cores with the unit bonds explicit),
and exposes the hot-path operations:

    inner / norm      <- TensorNetwork.inner / norm        pytens/algs.py:585-594
    right_orth        <- tt_right_orth                     pytens/algs.py:1654-1704
    round             <- tt_svd_round (relative eps)       pytens/algs.py:1841-1903
    dense             <- TensorNetwork.contract (chain)    pytens/algs.py:469-485

PyTorch is used for buffer ownership and streams only; all arithmetic happens in
libttb200.so.
"""

from __future__ import annotations

import math
import threading
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import TTDescriptor, check

_WORKSPACES: dict = {}


def workspace(nbytes: int, device: torch.device, slot: str = "main") -> torch.Tensor:
    """A cached, grow-only uint8 scratch buffer per (device, slot, host thread, current stream): calls from several
    threads or on several streams never share scratch memory."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), slot, threading.get_ident(),
           torch.cuda.current_stream().cuda_stream)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _WORKSPACES.pop(key, None)
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


_COPY_STREAM = None  # side stream of the streamed (copy-overlapped) inner product

_PACK_PAD = np.zeros(1, dtype=np.float64)
_PACK_BUFFERS: dict = {}


def _pack_buffers(count: int):
    """Grow-only (pinned host, device) float64 staging pair of inner_host_packed, per (device, host thread, stream)."""
    key = (torch.cuda.current_device(), threading.get_ident(), torch.cuda.current_stream().cuda_stream)
    pair = _PACK_BUFFERS.get(key)
    if pair is None or pair[0].numel() < count:
        cap = max(int(count), 1 << 16)
        pair = (torch.empty(cap, dtype=torch.float64).pin_memory(), torch.empty(cap, dtype=torch.float64, device="cuda"))
        _PACK_BUFFERS[key] = pair
    return pair


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("tensor_networks_b200 needs a CUDA device (no CPU fallback)")


class TensorTrain:
    """A tensor train resident in HBM.  See the module docstring."""

    def __init__(self, cores: Sequence[torch.Tensor]):
        _require_cuda()
        cores = list(cores)
        if not cores:
            raise ValueError("a TensorTrain needs at least one core")
        for k, c in enumerate(cores):
            if c.dtype != torch.float64 or not c.is_cuda or c.dim() != 3:
                raise ValueError(f"core {k}: need a 3-d CUDA float64 tensor, got {c.dtype} {tuple(c.shape)}")
            if not c.is_contiguous():
                cores[k] = c.contiguous()
        if cores[0].shape[0] != 1 or cores[-1].shape[2] != 1:
            raise AssertionError("boundary bond ranks must be 1")
        for k in range(len(cores) - 1):
            if cores[k].shape[2] != cores[k + 1].shape[0]:
                raise AssertionError(
                    f"bond {k}: {tuple(cores[k].shape)} does not chain with {tuple(cores[k + 1].shape)}"
                )
        self.cores: List[torch.Tensor] = cores

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_cores(cls, cores: Sequence[np.ndarray], device="cuda", pinned: bool = False) -> "TensorTrain":
        """Upload host cores.  Accepts the reference's shapes (2-d first/last core)."""
        _require_cuda()
        d = len(cores)
        host = []
        for k, c in enumerate(cores):
            a = np.ascontiguousarray(np.asarray(c, dtype=np.float64))
            if a.ndim == 2 and d == 1:
                raise ValueError("a 1-core TT must be given as (1, n, 1)")
            if a.ndim == 2 and k == 0:
                a = a.reshape(1, a.shape[0], a.shape[1])
            elif a.ndim == 2 and k == d - 1:
                a = a.reshape(a.shape[0], a.shape[1], 1)
            elif a.ndim != 3:
                raise ValueError(f"core {k} has unsupported shape {a.shape}")
            host.append(a)
        if not pinned and sum(a.nbytes for a in host) >= (32 << 20):
            # large pageable upload: staged through the pinned ring by the library's host threads
            import ctypes

            L = _lib.lib()
            out = [torch.empty(a.shape, dtype=torch.float64, device=device) for a in host]
            dst = (ctypes.c_void_p * d)(*[int(t.data_ptr()) for t in out])
            src = (ctypes.c_void_p * d)(*[int(a.ctypes.data) for a in host])
            nb = (ctypes.c_size_t * d)(*[int(a.nbytes) for a in host])
            check(L.ttb_h2d_staged(dst, src, nb, d, _stream_ptr()))
            torch.cuda.current_stream().synchronize()  # the numpy buffers may go away after this call
            return cls(out)
        out = []
        for a in host:
            t = torch.from_numpy(a)
            if pinned:
                t = t.pin_memory()
            out.append(t.to(device, non_blocking=pinned))
        return cls(out)

    @classmethod
    def from_network(cls, tn, device="cuda") -> "TensorTrain":
        """From a pytens-style TensorNetwork whose nodes are 0..d-1 in a chain."""
        d = len(tn.network.nodes)
        vals = []
        for k in range(d):
            if k not in tn.network.nodes:
                raise ValueError("TT-shaped networks must have integer nodes 0..d-1 (pytens/algs.py:1870-1885)")
            vals.append(tn.network.nodes[k]["tensor"].value)
        return cls.from_cores(vals, device=device)

    @classmethod
    def rand(
        cls,
        shape: Sequence[int],
        ranks: Sequence[int],
        seed: Optional[int] = None,
        scaled: bool = True,
        device="cuda",
    ) -> "TensorTrain":
        """Random TT generated on the device (layout of rand_tt, pytens/algs.py:1180-1218).

        `scaled` multiplies core k by (n_k r_k)^(-1/2) so that ||X|| = O(1).
        """
        _require_cuda()
        d = len(shape)
        assert len(ranks) + 1 == d
        r = [1] + [int(x) for x in ranks] + [1]
        gen = torch.Generator(device=device)
        if seed is not None:
            gen.manual_seed(int(seed))
        cores = []
        for k in range(d):
            c = torch.randn((r[k], int(shape[k]), r[k + 1]), dtype=torch.float64, device=device, generator=gen)
            if scaled:
                c *= 1.0 / math.sqrt(shape[k] * r[k + 1])
            cores.append(c)
        return cls(cores)

    @classmethod
    def from_dense(cls, dense, eps: float, max_rank: Optional[int] = None) -> "TensorTrain":
        """TT-SVD of a dense tensor (numpy array or CUDA tensor) with relative accuracy eps.

        Sequential reshape-and-truncate with delta = eps / sqrt(d-1) * ||X||_F -- the
        composition TensorNetwork.svd + merge of the reference (pytens/algs.py:633-702,
        :735-761; delta_svd pytens/utils.py:19-100, delta formula :53).
        """
        import ctypes

        _require_cuda()
        L = _lib.lib()
        if isinstance(dense, np.ndarray):
            dense = torch.from_numpy(np.ascontiguousarray(dense, dtype=np.float64)).cuda()
        if dense.dtype != torch.float64 or not dense.is_cuda:
            raise ValueError("from_dense needs a float64 array / CUDA tensor")
        dense = dense.contiguous()
        shape = [int(x) for x in dense.shape]
        d = len(shape)
        shp = (ctypes.c_int64 * d)(*shape)
        ws = workspace(L.ttb_ttsvd_workspace_bytes(d, shp), dense.device)
        # worst-case size of all cores: bond k is at most min(prod n_{<=k}, prod n_{>k}) (and max_rank)
        total, left, need, r_prev = dense.numel(), 1, 0, 1
        for k in range(d):
            left *= shape[k]
            r_next = 1 if k == d - 1 else min(left, total // left, 8192)
            m = r_prev * shape[k]
            need += m * (1 if k == d - 1 else min(m, total // left))  # U is written as m x min(m, c)
            if max_rank:
                r_next = min(r_next, int(max_rank))
            r_prev = r_next
        arena = torch.empty(max(need, 1), dtype=torch.float64, device=dense.device)
        ranks = (ctypes.c_int64 * (d + 1))()
        delta = ctypes.c_double(0.0)
        check(
            L.ttb_ttsvd_f64(
                dense.data_ptr(), d, shp, float(eps), int(max_rank) if max_rank else 0, arena.data_ptr(),
                arena.numel(), ranks, ctypes.byref(delta), ws.data_ptr(), ws.numel(), _stream_ptr(),
            )
        )
        cores, off = [], 0
        for k in range(d):
            rl, n, rr = int(ranks[k]), shape[k], int(ranks[k + 1])
            cores.append(arena[off : off + rl * n * rr].view(rl, n, rr).clone())
            off += rl * n * rr
        tt = cls(cores)
        tt.last_ttsvd = {"delta": float(delta.value)}
        return tt

    def clone(self) -> "TensorTrain":
        return TensorTrain([c.clone() for c in self.cores])

    # ------------------------------------------------------------------ accessors
    @property
    def d(self) -> int:
        return len(self.cores)

    def dim(self) -> int:
        return len(self.cores)

    def shape(self) -> List[int]:
        return [int(c.shape[1]) for c in self.cores]

    def ranks(self) -> List[int]:
        return [int(c.shape[2]) for c in self.cores[:-1]]

    def bond_ranks(self) -> List[int]:
        return [1] + self.ranks() + [1]

    @property
    def device(self) -> torch.device:
        return self.cores[0].device

    def nbytes(self) -> int:
        return sum(c.numel() * 8 for c in self.cores)

    def to_cores(self) -> List[np.ndarray]:
        return [c.detach().cpu().numpy() for c in self.cores]

    def descriptor(self) -> TTDescriptor:
        return TTDescriptor(self.shape(), self.bond_ranks(), [c.data_ptr() for c in self.cores])

    def scale(self, alpha: float) -> "TensorTrain":
        """In place, on the first core -- TensorNetwork.scale, pytens/algs.py:578-583."""
        from . import dense

        dense.scal(self.cores[0], float(alpha))
        return self

    def __add__(self, other: "TensorTrain") -> "TensorTrain":
        """Formal sum by block-diagonal rank growth (pytens/algs.py:1339-1353, :308-344): block
        placements into zero-filled cores (`ttb_fill_f64` + `ttb_strided_copy_f64`)."""
        from . import dense

        if self.shape() != other.shape():
            raise AssertionError("TT sum needs equal mode sizes")
        d = self.d
        out = []
        for k, (a, b) in enumerate(zip(self.cores, other.cores)):
            ra, n, rb = (int(x) for x in a.shape)
            sa, _, sb = (int(x) for x in b.shape)
            if d == 1:
                c = dense.empty(a.shape, a.device)
                dense.place(c, a, (0, 0, 0))
                dense.axpy(1.0, b, c)
            elif k == 0:
                c = dense.empty((1, n, rb + sb), a.device)
                dense.place(c, a, (0, 0, 0))
                dense.place(c, b, (0, 0, rb))
            elif k == d - 1:
                c = dense.empty((ra + sa, n, 1), a.device)
                dense.place(c, a, (0, 0, 0))
                dense.place(c, b, (ra, 0, 0))
            else:
                c = dense.zeros((ra + sa, n, rb + sb), a.device)
                dense.place(c, a, (0, 0, 0))
                dense.place(c, b, (ra, 0, rb))
            out.append(c)
        return TensorTrain(out)

    # ------------------------------------------------------------------ hot path
    @staticmethod
    def inner_streamed(host_a: Sequence[torch.Tensor], host_b: Sequence[torch.Tensor],
                       dev_a: Optional["TensorTrain"] = None, dev_b: Optional["TensorTrain"] = None) -> torch.Tensor:
        """<A, B> for two trains whose cores are PINNED HOST tensors (3-d float64, unit boundary bonds).

        The host->device copies run on a side stream (copy engines) while the persistent sweep kernel
        already runs and waits, core by core, for the data, so the transfer overlaps the contraction
        (`ttb_inner_streamed_f64`).  `dev_a` / `dev_b` may supply device trains of the same shapes to
        receive the copies (re-used across calls); otherwise they are allocated.  Returns a 0-d CUDA
        tensor; asynchronous with respect to the host like `inner_dev`."""
        import ctypes

        _require_cuda()
        L = _lib.lib()
        for h in list(host_a) + list(host_b):
            if h.dtype != torch.float64 or h.dim() != 3 or h.is_cuda or not h.is_pinned() or not h.is_contiguous():
                raise ValueError("inner_streamed needs contiguous pinned host float64 cores of shape (r, n, r')")
        if dev_a is None:
            dev_a = TensorTrain([torch.empty(h.shape, dtype=torch.float64, device="cuda") for h in host_a])
        if dev_b is None:
            dev_b = TensorTrain([torch.empty(h.shape, dtype=torch.float64, device="cuda") for h in host_b])
        if [tuple(c.shape) for c in dev_a.cores] != [tuple(h.shape) for h in host_a] or \
                [tuple(c.shape) for c in dev_b.cores] != [tuple(h.shape) for h in host_b]:
            raise AssertionError("inner_streamed: device trains do not match the host cores")
        if dev_a.shape() != dev_b.shape():
            raise AssertionError("inner: free indices (mode sizes) differ")
        da, db = dev_a.descriptor(), dev_b.descriptor()
        d = dev_a.d
        pa = (ctypes.c_void_p * d)(*[int(h.data_ptr()) for h in host_a])
        pb = (ctypes.c_void_p * d)(*[int(h.data_ptr()) for h in host_b])
        ws = workspace(L.ttb_inner_streamed_workspace_bytes(da.ref(), db.ref()), dev_a.device)
        out = torch.empty((), dtype=torch.float64, device=dev_a.device)
        global _COPY_STREAM
        if _COPY_STREAM is None:
            _COPY_STREAM = torch.cuda.Stream()
        check(L.ttb_inner_streamed_f64(da.ref(), db.ref(), pa, pb, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _stream_ptr(), _COPY_STREAM.cuda_stream))
        # the device trains are written by the copy stream: keep them alive until the compute stream is done
        out._ttb_keepalive = (dev_a, dev_b, host_a, host_b)
        return out

    @staticmethod
    def inner_host(host_a: Sequence[np.ndarray], host_b: Sequence[np.ndarray]) -> np.ndarray:
        """<A, B> for two trains whose cores are ordinary (pageable) numpy arrays -- what a pytens caller
        holds (`TensorNetwork.inner`, pytens/algs.py:585-587).  The reference's core shapes are accepted
        (2-d first / last core).  The persistent sweep kernel starts at once; host threads stage the cores
        through a pinned ring to the copy engines underneath it (`ttb_inner_streamed_f64`, csrc/staging.cu),
        so the transfer overlaps the contraction.  Returns a 0-d float64 array; synchronous."""
        import ctypes

        _require_cuda()
        L = _lib.lib()
        if len(host_a) != len(host_b):
            raise AssertionError("inner: operands have different numbers of cores")
        d = len(host_a)

        def prep(cores):
            out = []
            for k, c in enumerate(cores):
                a = np.ascontiguousarray(c, dtype=np.float64)
                if a.ndim == 2 and k == 0 and d > 1:
                    a = a.reshape(1, a.shape[0], a.shape[1])
                elif a.ndim == 2 and k == d - 1:
                    a = a.reshape(a.shape[0], a.shape[1], 1)
                elif a.ndim == 1 and d == 1:
                    a = a.reshape(1, -1, 1)
                elif a.ndim != 3:
                    raise ValueError(f"core {k} has unsupported shape {a.shape}")
                out.append(a)
            return out

        ha, hb = prep(host_a), prep(host_b)
        dev_a = TensorTrain([torch.empty(h.shape, dtype=torch.float64, device="cuda") for h in ha])
        dev_b = TensorTrain([torch.empty(h.shape, dtype=torch.float64, device="cuda") for h in hb])
        if dev_a.shape() != dev_b.shape():
            raise AssertionError("inner: free indices (mode sizes) differ")
        da, db = dev_a.descriptor(), dev_b.descriptor()
        pa = (ctypes.c_void_p * d)(*[int(h.ctypes.data) for h in ha])
        pb = (ctypes.c_void_p * d)(*[int(h.ctypes.data) for h in hb])
        ws = workspace(L.ttb_inner_streamed_workspace_bytes(da.ref(), db.ref()), dev_a.device)
        out = torch.empty((), dtype=torch.float64, device=dev_a.device)
        global _COPY_STREAM
        if _COPY_STREAM is None:
            _COPY_STREAM = torch.cuda.Stream()
        check(L.ttb_inner_streamed_f64(da.ref(), db.ref(), pa, pb, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _stream_ptr(), _COPY_STREAM.cuda_stream))
        val = out.item()  # synchronises the compute stream; ha / hb / dev_* stay alive until here
        _COPY_STREAM.synchronize()
        if val != val:
            raise RuntimeError("streamed inner product: a core did not arrive within the time-out "
                               "(or an operand contains NaN)")
        return np.asarray(val, dtype=np.float64)

    @staticmethod
    def inner_host_packed(host_a: Sequence[np.ndarray], host_b: Sequence[np.ndarray]) -> np.ndarray:
        """<A, B> for two SMALL trains held as numpy cores (the reference's shapes: 2-d first / last core):
        every core of both operands is packed into one pinned staging buffer (one `np.concatenate`), moved by ONE
        host-to-device copy, and the descriptors point into that buffer -- no per-core tensor, copy or allocation
        (the per-core upload cost 34 us per core: 22 ms for the d = 640 point of examples/inner_product_scaling.py).
        Returns a 0-d float64 array; synchronous."""
        _require_cuda()
        L = _lib.lib()
        d = len(host_a)
        if d != len(host_b) or d < 2:
            raise AssertionError("inner: operands have different numbers of cores")
        pieces, offs, shapes, total = [], [], [], 0
        pad = _PACK_PAD
        for cores in (host_a, host_b):
            for k, c in enumerate(cores):
                a = c if (c.dtype == np.float64 and c.flags.c_contiguous) else np.ascontiguousarray(c, dtype=np.float64)
                if a.ndim == 2 and k == 0:
                    shp = (1, a.shape[0], a.shape[1])
                elif a.ndim == 2 and k == d - 1:
                    shp = (a.shape[0], a.shape[1], 1)
                elif a.ndim == 3:
                    shp = a.shape
                else:
                    raise ValueError(f"core {k} has unsupported shape {a.shape}")
                pieces.append(a.reshape(-1))
                offs.append(total)
                shapes.append(shp)
                total += a.size
                if total & 1:  # keep every core 16-byte aligned (vectorised copies, TMA boxes)
                    pieces.append(pad)
                    total += 1
        sa, sb = shapes[:d], shapes[d:]
        na, nb = [s[1] for s in sa], [s[1] for s in sb]
        if na != nb:
            raise AssertionError("inner: free indices (mode sizes) differ")
        ra = [sa[0][0]] + [s[2] for s in sa]
        rb = [sb[0][0]] + [s[2] for s in sb]
        for k in range(d - 1):
            if sa[k][2] != sa[k + 1][0] or sb[k][2] != sb[k + 1][0]:
                raise AssertionError(f"inner: bond {k + 1} has inconsistent ranks")
        pin, dev = _pack_buffers(total)
        np.concatenate(pieces, out=pin.numpy()[:total])
        dev[:total].copy_(pin[:total], non_blocking=True)
        base = dev.data_ptr()
        da = TTDescriptor(na, ra, [base + 8 * o for o in offs[:d]])
        db = TTDescriptor(nb, rb, [base + 8 * o for o in offs[d:]])
        ws = workspace(L.ttb_inner_workspace_bytes(da.ref(), db.ref()), dev.device)
        out = torch.empty((), dtype=torch.float64, device=dev.device)
        check(L.ttb_inner_f64(da.ref(), db.ref(), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
        return np.asarray(out.item(), dtype=np.float64)  # synchronises: the staging buffers are free again

    def inner_dev(self, other: "TensorTrain", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """<self, other> as a 0-d CUDA tensor; no host synchronisation."""
        L = _lib.lib()
        da, db = self.descriptor(), other.descriptor()
        if self.shape() != other.shape():
            raise AssertionError("inner: free indices (mode sizes) differ")
        nbytes = L.ttb_inner_workspace_bytes(da.ref(), db.ref())
        ws = workspace(nbytes, self.device)
        if out is None:
            out = torch.empty((), dtype=torch.float64, device=self.device)
        check(L.ttb_inner_f64(da.ref(), db.ref(), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
        return out

    def inner(self, other: "TensorTrain") -> np.ndarray:
        """0-d float64 ndarray, like TensorNetwork.inner (pytens/algs.py:585-587)."""
        return np.asarray(self.inner_dev(other).item(), dtype=np.float64)

    def norm(self) -> float:
        """sqrt(|<X, X>|) -- TensorNetwork.norm, pytens/algs.py:589-594."""
        val = float(self.inner_dev(self).item())
        return float(np.sqrt(np.abs(val)))

    def right_orth(self, node: int) -> "TensorTrain":
        """One RQ step on core `node`, in place -- tt_right_orth, pytens/algs.py:1654-1704.

        Core `node` gets orthonormal rows (horizontal unfolding) and R^T is pushed into
        core node-1.  Interior cores keep their rank (zero-padded when n*r_right < r_left),
        the last core shrinks to min(r, n), exactly like the reference.
        """
        import ctypes

        L = _lib.lib()
        if not 1 <= node <= self.d - 1:
            raise ValueError("right_orth: node must be in 1..d-1")
        desc = self.descriptor()
        nbytes = L.ttb_right_orth_workspace_bytes(desc.ref(), node)
        ws = workspace(nbytes, self.device)
        new_rank = ctypes.c_int64(0)
        check(L.ttb_right_orth_f64(desc.ref(), node, ctypes.byref(new_rank), ws.data_ptr(), ws.numel(), _stream_ptr()))
        c_new = int(new_rank.value)
        ck, cp = self.cores[node], self.cores[node - 1]
        if c_new != ck.shape[0]:
            self.cores[node] = ck.reshape(-1)[: c_new * ck.shape[1] * ck.shape[2]].view(c_new, ck.shape[1], ck.shape[2])
            self.cores[node - 1] = cp.reshape(-1)[: cp.shape[0] * cp.shape[1] * c_new].view(cp.shape[0], cp.shape[1], c_new)
        return self

    def round(self, eps: float, max_rank: Optional[int] = None) -> "TensorTrain":
        """Round in place with relative accuracy eps -- tt_svd_round, pytens/algs.py:1841-1903.

        RQ pass then left-to-right delta-truncated SVD sweep with
        delta = eps / sqrt(d-1) * ||X||_F.  `max_rank` (not in the reference) caps every
        bond: rank = min(rank_eps, max_rank).  Returns self; `self.last_round` holds
        {"delta", "svds", "jacobi_sweeps", "not_converged"}.
        """
        import ctypes

        L = _lib.lib()
        d = self.d
        desc = self.descriptor()
        nbytes = L.ttb_round_workspace_bytes(desc.ref())
        ws = workspace(nbytes, self.device)
        ranks = (ctypes.c_int64 * (d + 1))()
        delta = ctypes.c_double(0.0)
        stats = (ctypes.c_int32 * 5)()
        check(
            L.ttb_round_f64(
                desc.ref(), float(eps), int(max_rank) if max_rank else 0, ranks, ctypes.byref(delta), stats,
                ws.data_ptr(), ws.numel(), _stream_ptr(),
            )
        )
        new = []
        for k, c in enumerate(self.cores):
            rl, n, rr = int(ranks[k]), int(c.shape[1]), int(ranks[k + 1])
            new.append(c.reshape(-1)[: rl * n * rr].view(rl, n, rr))
        self.cores = new
        self.last_round = {
            "delta": float(delta.value),
            "svds": int(stats[0]),
            "jacobi_sweeps": int(stats[1]),
            "not_converged": int(stats[2]),
            "svds_certified": int(stats[3]),
            "bonds_deflated": int(stats[4]),
        }
        return self

    def gramsvd_round(self, eps: float) -> "TensorTrain":
        """Round in place by Gram SVD -- tt_gramsvd_round, pytens/algs.py:1771-1838 (see gramsvd.py)."""
        from .gramsvd import gramsvd_round

        return gramsvd_round(self, eps)

    def compact(self) -> "TensorTrain":
        """Re-own the cores (drop the slack left in the buffers by an in-place rounding)."""
        self.cores = [c.clone() for c in self.cores]
        return self

    def dense_dev(self) -> torch.Tensor:
        """The dense tensor (n_1, ..., n_d) on the device."""
        L = _lib.lib()
        desc = self.descriptor()
        nbytes = L.ttb_tt_to_dense_workspace_bytes(desc.ref())
        ws = workspace(nbytes, self.device)
        out = torch.empty(self.shape(), dtype=torch.float64, device=self.device)
        check(L.ttb_tt_to_dense_f64(desc.ref(), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
        return out

    def dense(self) -> np.ndarray:
        return self.dense_dev().cpu().numpy()
