"""TT operators, TT sums and TT-GMRES with the reference's signatures, computed on the device.

Mirror of `ttop_rank1` / `ttop_rank2` / `ttop_sum` / `tt_sum` / `ttop_sum_apply` / `ttop_apply` / `gmres`
(pytens/algs.py:2383-2793; exercised by tests/main_test.py:264-448).  They work on `TensorNetwork`
objects whose node values are numpy arrays or CUDA tensors.  The block-diagonal builders are block
placements (`ttb_strided_copy_f64` into a zero-filled core), the operator application is one DMMA GEMM
per core (`ttb_gemm_f64`) between permuted operands, and GMRES calls the rounding / inner-product
sweeps of this package per Arnoldi step; only the small Hessenberg least-squares problem runs on
the host, as in the reference.  (`solvers.py` holds the same algorithms on bare `TensorTrain`s.)
"""

from __future__ import annotations

import copy
from typing import Callable, List

import numpy as np

from . import dense
from .types import Index


def _net():
    from . import algs

    return algs


def _blockdiag(blocks: list, shape, offsets_of) -> "dense.torch.Tensor":
    out = dense.zeros(shape, blocks[0].device)
    for j, b in enumerate(blocks):
        dense.place(out, b, offsets_of(j, b))
    return out


def ttop_rank1(indices_in: List[Index], indices_out: List[Index], cores: List[np.ndarray],
               rank_name_prefix: str):
    """Rank-1 TT operator A_1 (x) ... (x) A_d with unit bonds -- pytens/algs.py:2383-2425.
    Node k: (rank, out, in, rank); first node (out, in, rank), last (rank, out, in)."""
    assert len(indices_in) == len(indices_out)
    A = _net()
    dim = len(indices_in)
    op = A.TensorNetwork()
    bonds = [Index(f"{rank_name_prefix}_r{k + 1}", 1) for k in range(dim)]
    for k in range(dim):
        c = cores[k]
        if k == 0:
            op.add_node(0, A.Tensor(c[:, :, None], [indices_out[0], indices_in[0], bonds[0]]))
        elif k < dim - 1:
            op.add_node(k, A.Tensor(c[None, :, :, None], [bonds[k - 1], indices_out[k], indices_in[k], bonds[k]]))
        else:
            op.add_node(k, A.Tensor(c[None, :, :], [bonds[k - 1], indices_out[k], indices_in[k]]))
        if k > 0:
            op.add_edge(k - 1, k)
    return op


def ttop_sum(indices_in: List[Index], indices_out: List[Index], cores: List[List[np.ndarray]],
             rank_name_prefix: str):
    """Sum of len(cores) rank-1 operators as one TT operator with diagonal bond structure --
    pytens/algs.py:2479-2532 (cores[j][k]: factor k of term j)."""
    assert len(indices_in) == len(indices_out)
    A = _net()
    dim, ns = len(indices_in), len(cores)
    op = A.TensorNetwork()
    bonds = [Index(f"{rank_name_prefix}_r{k + 1}", ns) for k in range(dim)]
    for k in range(dim):
        no, ni = indices_out[k].size, indices_in[k].size
        mats = [dense.as_dev(cores[j][k]) for j in range(ns)]
        host = not dense.is_dev(cores[0][k])
        if k == 0:
            val = _blockdiag([m.view(no, ni, 1) for m in mats], (no, ni, ns), lambda j, b: (0, 0, j))
            inds = [indices_out[0], indices_in[0], bonds[0]]
        elif k < dim - 1:
            val = _blockdiag([m.view(1, no, ni, 1) for m in mats], (ns, no, ni, ns), lambda j, b: (j, 0, 0, j))
            inds = [bonds[k - 1], indices_out[k], indices_in[k], bonds[k]]
        else:
            val = _blockdiag([m.view(1, no, ni) for m in mats], (ns, no, ni), lambda j, b: (j, 0, 0))
            inds = [bonds[k - 1], indices_out[k], indices_in[k]]
        op.add_node(k, A.Tensor(dense.to_host(val) if host else val, inds))
        if k > 0:
            op.add_edge(k - 1, k)
    return op


def ttop_rank2(indices_in: List[Index], indices_out: List[Index], cores_r1: List[np.ndarray],
               cores_r2: List[np.ndarray], rank_name_prefix: str):
    """Sum of two rank-1 operators -- pytens/algs.py:2428-2476."""
    return ttop_sum(indices_in, indices_out, [cores_r1, cores_r2], rank_name_prefix)


def tt_sum(tt_in: list):
    """Sum of several TTs of equal mode sizes by block-diagonal bond growth -- pytens/algs.py:2535-2585.
    Bond indices are named rank_k, free indices keep the names of the first summand."""
    A = _net()
    out = A.TensorNetwork()
    first = tt_in[0]
    dim = first.dim()
    for ii, node in enumerate(first.network.nodes):
        inds = first.network.nodes[node]["tensor"].indices
        raw = [tt.value(node) for tt in tt_in]
        host = not dense.is_dev(raw[0])
        vals = [dense.as_dev(v) for v in raw]
        if ii == 0:
            total = sum(int(v.shape[1]) for v in vals)
            offs = np.cumsum([0] + [int(v.shape[1]) for v in vals])
            new = _blockdiag(vals, (int(vals[0].shape[0]), total), lambda j, b: (0, int(offs[j])))
            new_inds = [Index(inds[0].name, inds[0].size), Index("rank_0", total)]
        elif ii == dim - 1:
            total = sum(int(v.shape[0]) for v in vals)
            offs = np.cumsum([0] + [int(v.shape[0]) for v in vals])
            new = _blockdiag(vals, (total, int(vals[0].shape[1])), lambda j, b: (int(offs[j]), 0))
            new_inds = [Index(f"rank_{ii - 1}", total), Index(inds[1].name, inds[1].size)]
        else:
            lo = np.cumsum([0] + [int(v.shape[0]) for v in vals])
            ro = np.cumsum([0] + [int(v.shape[2]) for v in vals])
            new = _blockdiag(vals, (int(lo[-1]), int(vals[0].shape[1]), int(ro[-1])),
                             lambda j, b: (int(lo[j]), 0, int(ro[j])))
            new_inds = [Index(f"rank_{ii - 1}", int(lo[-1])), Index(inds[1].name, inds[1].size),
                        Index(f"rank_{ii}", int(ro[-1]))]
        out.add_node(ii, A.Tensor(dense.to_host(new) if host else new, new_inds))
        if ii > 0:
            out.add_edge(ii - 1, ii)
    return out


def ttop_sum_apply(tt_in, indices_in: List[Index], indices_out: List[Index],
                   cores: List[List[Callable]], rank_name_prefix: str):
    """Apply a sum of rank-1 operators given as per-core callables -- pytens/algs.py:2588-2660.
    cores[j][k](v) maps core k of the input (numpy array or CUDA tensor, as stored) to the core of
    term j; the results are placed block-diagonally."""
    assert len(indices_in) == len(indices_out)
    A = _net()
    dim, ns = len(indices_in), len(cores)
    out = A.TensorNetwork()
    nodes = list(tt_in.network.nodes())
    bonds: List[Index] = []
    for ii, node in enumerate(nodes):
        v = tt_in.value(node)
        host = not dense.is_dev(v)
        terms = [dense.as_dev(cores[jj][ii](v)) for jj in range(ns)]
        no = indices_out[ii].size
        if ii == 0:
            terms = [t.reshape(no, -1) for t in terms]
            offs = np.cumsum([0] + [int(t.shape[1]) for t in terms])
            bonds.append(Index(f"{rank_name_prefix}_r1", ns * int(v.shape[1])))
            new = _blockdiag(terms, (no, ns * int(v.shape[1])), lambda j, b: (0, int(offs[j])))
            inds = [indices_out[0], bonds[0]]
        elif ii < dim - 1:
            terms = [t.reshape(t.shape[0], t.shape[1], t.shape[2]) for t in terms]
            lo = np.cumsum([0] + [int(t.shape[0]) for t in terms])
            ro = np.cumsum([0] + [int(t.shape[2]) for t in terms])
            bonds.append(Index(f"{rank_name_prefix}_r{ii + 1}", int(v.shape[2]) * ns))
            new = _blockdiag(terms, (ns * int(v.shape[0]), no, ns * int(v.shape[2])),
                             lambda j, b: (int(lo[j]), 0, int(ro[j])))
            inds = [bonds[ii - 1], indices_out[ii], bonds[ii]]
        else:
            lo = np.cumsum([0] + [int(t.shape[0]) for t in terms])
            new = _blockdiag(terms, (ns * int(v.shape[0]), no), lambda j, b: (int(lo[j]), 0))
            inds = [bonds[ii - 1], indices_out[ii]]
        out.add_node(ii, A.Tensor(dense.to_host(new) if host else new, inds))
        if ii > 0:
            out.add_edge(ii - 1, ii)
    return out


def ttop_apply(ttop, tt_in):
    """y = op(x) core by core -- pytens/algs.py:2662-2697: interior 'ijkl,mkp->mijpl' reshaped to
    (m i, j, p l) (TT rank major, operator rank minor), first 'ijk,jl->ilk' -> (n, l k), last
    'ijk,mk->mij' -> (m i, j).  Each core is one GEMM between permuted operands.  Returns a copy of
    tt_in with the new cores (same index names, resized)."""
    tt = copy.deepcopy(tt_in)
    dim = tt.dim()
    for ii, (node_op, node_tt) in enumerate(zip(ttop.network.nodes(), tt.network.nodes())):
        raw = tt.network.nodes[node_tt]["tensor"].value
        host = not dense.is_dev(raw)
        op = dense.as_dev(ttop.network.nodes[node_op]["tensor"].value)
        v = dense.as_dev(raw)
        if ii == 0:
            c, _ = dense.contract(op, "ijk", v, "jl")  # -> i k l
            new = dense.permute(c, [0, 2, 1]).reshape(int(v.shape[0]), -1)
        elif ii < dim - 1:
            c, _ = dense.contract(op, "ijkl", v, "mkp")  # -> i j l m p
            c = dense.permute(c, [3, 0, 1, 4, 2])  # m i j p l
            s = [int(x) for x in c.shape]
            new = c.reshape(s[0] * s[1], s[2], s[3] * s[4])
        else:
            c, _ = dense.contract(op, "ijk", v, "mk")  # -> i j m
            c = dense.permute(c, [2, 0, 1])
            new = c.reshape(int(c.shape[0]) * int(c.shape[1]), -1)
        t = tt.network.nodes[node_tt]["tensor"]
        tt.network.nodes[node_tt]["tensor"] = t.update_val_size(dense.to_host(new) if host else new)
    return tt


def gmres(op, rhs, x0, eps: float = 1e-5, round_eps: float = 1e-10, maxiter: int = 100):
    """TT-GMRES, line by line after pytens/algs.py:2701-2793: modified Gram-Schmidt against the Krylov
    basis with `inner`, `tt_svd_round(w, round_eps)` after the operator and after the orthogonalisation,
    a dense least-squares solve of the Hessenberg system on the host; stops when its residual is below
    eps.  Returns (x, ||rhs - op(x)||).  All TT arithmetic runs on the device."""
    A = _net()
    r0 = rhs + op(x0).scale(-1.0)
    r0 = A.tt_svd_round(r0, round_eps)
    beta = r0.norm()
    r0.scale(1.0 / beta)
    v = [r0]
    y = []
    H = None
    for jj in range(maxiter):
        w = op(v[-1])
        w = A.tt_svd_round(w, round_eps)
        if H is None:
            H = np.zeros((jj + 2, jj + 1))
        else:
            m, n = H.shape
            grown = np.zeros((m + 1, n + 1))
            grown[:m, :n] = H
            H = grown
        for ii in range(jj + 1):
            H[ii, jj] = w.inner(v[ii])
            vv = copy.deepcopy(v[ii])
            vv.scale(-H[ii, jj])
            w = w + vv
        w = A.tt_svd_round(w, round_eps)
        H[jj + 1, jj] = w.norm()
        v.append(w.scale(1.0 / H[jj + 1, jj]))
        e = np.zeros(H.shape[0])
        e[0] = beta
        yy, resid, _, _ = np.linalg.lstsq(H, e, rcond=None)
        y.append(yy)
        if resid.size > 0 and abs(float(resid[0])) < eps:
            break
    x = copy.deepcopy(x0)
    for vv, yy in zip(v, y[-1]):
        x = x + vv.scale(yy)
    x = A.tt_svd_round(x, round_eps)
    r0 = rhs + op(x).scale(-1.0)
    return x, r0.norm()
