"""Host-side mirror of the `pytens.algs` surface for the tensor-train hot path.

The reference has no plugin/FFI boundary: its public API *is* the Python functions of
`pytens/algs.py`.  This module keeps those names, argument meanings and mutation /
error behaviour for the TT core-sweep path and routes the arithmetic to the CUDA
library through `TensorTrain` (ctypes -> libttb200.so):

    Tensor, TensorNetwork (TT-shaped networks)      pytens/algs.py:46-345, :363-631
    TensorNetwork.inner / norm / scale / rand_tt     :585-594, :578-583, :1180-1218
    tt_right_orth(tn, node)                          :1654-1704
    tt_svd_round(tn, eps)                            :1841-1903
    tt_gramsvd_round(tn, eps), gram_eig_and_svd,
    eps_to_rank (from .gramsvd)                      :1707-1838
    delta_svd (re-exported from .utils)              pytens/utils.py:19-100
    tt_svd(dense, eps)   [composition, no single reference function: SURVEY 3.3]

Everything else in pytens (general trees, cross approximation, structure search,
plotting) is out of scope and raises NotImplementedError here rather than falling
back to a CPU implementation.  The functions also accept the reference's own
`pytens.TensorNetwork` objects (duck-typed on `.network.nodes[k]["tensor"]`), which is
how a pytens installation would adopt this path -- see INTEGRATION.md.
"""

from __future__ import annotations

import copy
from collections import Counter
from dataclasses import dataclass
from typing import List, Optional, Sequence

import networkx as nx
import numpy as np

from .types import Index, IntOrStr, NodeName, SVDConfig  # noqa: F401
from .utils import TruncSVD, delta_svd  # noqa: F401
from .gramsvd import eps_to_rank, gram_eig_and_svd  # noqa: F401
from .tt import TensorTrain
from .solvers import TTOperator, gmres, ttop_apply, ttop_rank1  # noqa: F401  (device-resident TT-GMRES)

__all__ = [
    "Index", "SVDConfig", "Tensor", "TensorNetwork", "TensorTrain", "TruncSVD",
    "delta_svd", "tt_right_orth", "tt_svd_round", "tt_gramsvd_round", "eps_to_rank", "gram_eig_and_svd", "tt_svd", "round",
    "TTOperator", "ttop_rank1", "ttop_apply", "gmres",
]


@dataclass
class Tensor:
    """value + indices (pytens/algs.py:46-78)."""

    value: np.ndarray
    indices: List[Index]

    def update_val_size(self, value: np.ndarray) -> "Tensor":
        """Rebind the value (no copy) and resize the indices -- pytens/algs.py:70-78."""
        assert value.ndim == len(self.indices), f"{value.shape}, {self.indices}"
        self.value = value
        for ii, index in enumerate(self.indices):
            self.indices[ii] = index.with_new_size(value.shape[ii])
        return self


def _is_chain(tn) -> bool:
    nodes = list(tn.network.nodes)
    return nodes == list(range(len(nodes)))


def _tt_cores(tn) -> List[np.ndarray]:
    """Cores of a TT-shaped network (int nodes 0..d-1 in a chain, pytens/algs.py:1188-1216)."""
    if not _is_chain(tn):
        raise NotImplementedError(
            "only TT-shaped networks (integer nodes 0..d-1, as built by rand_tt) are supported on the "
            "B200 path; general tensor networks are out of scope (DESIGN.md)"
        )
    d = len(tn.network.nodes)
    vals = [tn.network.nodes[k]["tensor"].value for k in range(d)]
    for k, v in enumerate(vals):
        want = 2 if (k == 0 or k == d - 1) else 3
        if d == 1:
            want = 1
        if v.ndim != want:
            raise NotImplementedError(f"node {k}: expected a {want}-d TT core, got shape {v.shape}")
    return vals


def _free_index(tn, k: int, d: int) -> Index:
    inds = tn.network.nodes[k]["tensor"].indices
    if d == 1:
        return inds[0]
    return inds[0] if k == 0 else inds[1]


def _write_back(tn, tt: TensorTrain) -> None:
    """Store the device cores into the network in the reference's shapes (in place)."""
    d = tt.d
    host = tt.to_cores()
    for k, c in enumerate(host):
        if k == 0 and d > 1:
            c = c.reshape(c.shape[1], c.shape[2])
        elif k == d - 1 and d > 1:
            c = c.reshape(c.shape[0], c.shape[1])
        tn.network.nodes[k]["tensor"].update_val_size(np.ascontiguousarray(c))


class TensorNetwork:
    """TT-shaped tensor network with the reference's container layout (pytens/algs.py:363-631):
    `self.network` is an `nx.Graph`, node k carries attribute "tensor" = Tensor(value, indices)."""

    def __init__(self) -> None:
        self.network = nx.Graph()

    # ---- container (pytens/algs.py:370-444, :574-583) ----
    def add_node(self, name: NodeName, tensor: Tensor) -> None:
        self.network.add_node(name, tensor=tensor)

    def node_tensor(self, node_name: NodeName) -> Tensor:
        return self.network.nodes[node_name]["tensor"]

    def set_node_tensor(self, node_name: NodeName, value: Tensor) -> None:
        self.network.nodes[node_name]["tensor"] = value

    def add_edge(self, name1: NodeName, name2: NodeName) -> None:
        self.network.add_edge(name1, name2)

    def value(self, node_name: NodeName) -> np.ndarray:
        return self.network.nodes[node_name]["tensor"].value

    def all_indices(self) -> Counter:
        indices = []
        for _, data in self.network.nodes(data=True):
            indices += data["tensor"].indices
        return Counter(indices)

    def free_indices(self) -> List[Index]:
        return [i for i, v in self.all_indices().items() if v == 1]

    def inner_indices(self) -> List[Index]:
        return [i for i, v in self.all_indices().items() if v > 1]

    def ranks(self) -> List[int]:
        return [r.size for r in self.inner_indices()]

    def shape(self) -> List[int]:
        return [i.size for i in self.free_indices()]

    def dim(self) -> int:
        return len(self.free_indices())

    def scale(self, scale_factor: float) -> "TensorNetwork":
        """Multiply the first node in place -- pytens/algs.py:578-583."""
        for _, data in self.network.nodes(data=True):
            data["tensor"].value *= scale_factor
            break
        return self

    # ---- hot path ----
    def inner(self, other: "TensorNetwork") -> np.ndarray:
        """<self, other> over the shared free indices; 0-d float64 array like the reference
        (pytens/algs.py:585-587).  Replaces attach() + contract() by the device sweep."""
        a, b = _tt_cores(self), _tt_cores(other)
        if len(a) != len(b):
            raise AssertionError("inner: networks have different numbers of nodes")
        d = len(a)
        for k in range(d):
            ia, ib = _free_index(self, k, d), _free_index(other, k, d)
            if ia != ib:  # attach() only contracts indices with equal (name, size), algs.py:534-557
                raise NotImplementedError(f"inner: free index of node {k} differs ({ia} vs {ib})")
        ta = TensorTrain.from_cores(a if d > 1 else [a[0].reshape(1, -1, 1)])
        tb = TensorTrain.from_cores(b if d > 1 else [b[0].reshape(1, -1, 1)])
        return ta.inner(tb)

    def norm(self) -> float:
        """sqrt(|<self, self>|) -- pytens/algs.py:589-594."""
        val = float(self.inner(self))
        return float(np.sqrt(np.abs(val)))

    def contract(self) -> Tensor:
        """Dense tensor of a chain (TensorNetwork.contract, pytens/algs.py:469-485), on the device."""
        cores = _tt_cores(self)
        d = len(cores)
        tt = TensorTrain.from_cores(cores if d > 1 else [cores[0].reshape(1, -1, 1)])
        return Tensor(tt.dense(), [_free_index(self, k, d) for k in range(d)])

    def to_tensor_train(self) -> TensorTrain:
        return TensorTrain.from_network(self)

    @staticmethod
    def from_tensor_train(tt: TensorTrain, names: Optional[Sequence[IntOrStr]] = None) -> "TensorNetwork":
        """Chain network with the reference's naming (rand_tt, pytens/algs.py:1180-1218)."""
        d = tt.d
        shape, ranks = tt.shape(), tt.ranks()
        names = list(names) if names is not None else [f"x{k}" for k in range(d)]
        free = [Index(names[k], shape[k]) for k in range(d)]
        bonds = [Index(f"r{k + 1}", ranks[k]) for k in range(d - 1)]
        host = tt.to_cores()
        tn = TensorNetwork()
        for k in range(d):
            c = host[k]
            if d == 1:
                tn.add_node(0, Tensor(c.reshape(-1), [free[0]]))
            elif k == 0:
                tn.add_node(0, Tensor(c.reshape(c.shape[1], c.shape[2]), [free[0], bonds[0]]))
            elif k == d - 1:
                tn.add_node(k, Tensor(c.reshape(c.shape[0], c.shape[1]), [bonds[-1], free[k]]))
            else:
                tn.add_node(k, Tensor(c, [bonds[k - 1], free[k], bonds[k]]))
            if k > 0:
                tn.add_edge(k - 1, k)
        return tn

    @staticmethod
    def rand_tt(indices: List[Index], ranks: List[int]) -> "TensorNetwork":
        """Random TT with np.random.randn cores, same draw order as the reference
        (pytens/algs.py:1180-1218), so seeded fixtures agree value for value."""
        dim = len(indices)
        assert len(ranks) + 1 == len(indices)
        tt = TensorNetwork()
        r = [Index("r1", ranks[0])]
        tt.add_node(0, Tensor(np.random.randn(indices[0].size, ranks[0]), [indices[0], r[0]]))
        core = 1
        for ii, index in enumerate(indices[1:-1]):
            r.append(Index(f"r{ii + 2}", ranks[ii + 1]))
            tt.add_node(core, Tensor(np.random.randn(ranks[ii], index.size, ranks[ii + 1]), [r[ii], index, r[ii + 1]]))
            core += 1
            tt.add_edge(ii, ii + 1)
        tt.add_node(dim - 1, Tensor(np.random.randn(ranks[-1], indices[-1].size), [r[-1], indices[-1]]))
        tt.add_edge(dim - 2, dim - 1)
        return tt

    def __add__(self, other: "TensorNetwork") -> "TensorNetwork":
        """TT sum by block-diagonal rank growth (pytens/algs.py:1339-1353): host-side fixture
        builder, no arithmetic."""
        a, b = _tt_cores(self), _tt_cores(other)
        d = len(a)
        out = copy.deepcopy(self)
        for k in range(d):
            x, y = a[k], b[k]
            if d == 1:
                v = x + y
            elif k == 0:
                v = np.concatenate([x, y], axis=1)
            elif k == d - 1:
                v = np.concatenate([x, y], axis=0)
            else:
                v = np.zeros((x.shape[0] + y.shape[0], x.shape[1], x.shape[2] + y.shape[2]))
                v[: x.shape[0], :, : x.shape[2]] = x
                v[x.shape[0]:, :, x.shape[2]:] = y
            out.network.nodes[k]["tensor"].update_val_size(v)
        return out

    # ---- out of scope ----
    def attach(self, *_a, **_k):
        raise NotImplementedError("attach() is bypassed on the B200 path; use inner()")

    def round(self, node_name, delta, visited=None):
        raise NotImplementedError(
            "general-tree TensorNetwork.round (pytens/algs.py:763-827) is out of scope; use tt_svd_round / round()"
        )


def tt_right_orth(tn, node: int):
    """Right-orthogonalise core `node`, in place; returns tn -- pytens/algs.py:1654-1704."""
    _tt_cores(tn)
    tt = TensorTrain.from_network(tn)
    tt.right_orth(node)
    _write_back(tn, tt)
    return tn


def tt_svd_round(tn, eps: float):
    """Round a TT with relative accuracy eps, in place; returns the same object
    (pytens/algs.py:1841-1903)."""
    _tt_cores(tn)
    tt = TensorTrain.from_network(tn)
    tt.round(eps)
    _write_back(tn, tt)
    return tn


def tt_gramsvd_round(tn, eps: float):
    """Gram-SVD rounding of a TT, in place; returns the same object (pytens/algs.py:1771-1838)."""
    from .gramsvd import gramsvd_round

    _tt_cores(tn)
    tt = TensorTrain.from_network(tn)
    gramsvd_round(tt, eps)
    _write_back(tn, tt)
    return tn


def round(tn, eps: float, max_rank: Optional[int] = None):  # noqa: A001 (name fixed by north_star)
    """tt_svd_round with an optional rank cap: rank = min(rank_eps, max_rank).  The reference
    has no max_rank in rounding; parity of the cap is unpinned."""
    _tt_cores(tn)
    tt = TensorTrain.from_network(tn)
    tt.round(eps, max_rank=max_rank)
    _write_back(tn, tt)
    return tn


def tt_svd(dense: np.ndarray, eps: float, max_rank: Optional[int] = None,
           names: Optional[Sequence[IntOrStr]] = None) -> TensorNetwork:
    """TT-SVD of a dense array with delta = eps / sqrt(d-1) * ||X||_F; the composition
    TensorNetwork.svd + merge of the reference (pytens/algs.py:633-702, :735-761)."""
    tt = TensorTrain.from_dense(np.asarray(dense, dtype=np.float64), eps, max_rank=max_rank)
    return TensorNetwork.from_tensor_train(tt, names)
