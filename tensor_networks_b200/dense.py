"""Node-level dense-tensor operations on the device, through the ttb200 C ABI.

The tensor-network functions of the reference above the TT sweeps -- `Tensor.svd`, `Tensor.qr`,
`Tensor.contract`, `Tensor.permute`, `Tensor.block_diagonal`, `Tensor.mult`
(pytens/algs.py:143-344) -- are numpy `permute_dims` / `reshape` / `einsum` / `linalg.qr` /
`delta_svd` calls on one or two node tensors.  Here each of them is a permute-to-matrix copy
(`ttb_strided_copy_f64`), one DMMA GEMM (`ttb_gemm_f64`), the row-space QR (`ttb_orth_rows_f64`) or
the Jacobi truncated SVD (`ttb_delta_svd_f64`).  PyTorch only owns the buffers: no torch arithmetic
is used (allocation with `torch.empty`, views with `.view` / `.reshape` on contiguous data).
"""

from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check
from .tt import _require_cuda, _stream_ptr, workspace


def _i64(vals: Sequence[int]):
    return (ctypes.c_int64 * max(len(vals), 1))(*[int(v) for v in vals])


def is_dev(x) -> bool:
    return isinstance(x, torch.Tensor)


def as_dev(x) -> torch.Tensor:
    """A contiguous CUDA float64 tensor holding `x` (numpy array or torch tensor)."""
    _require_cuda()
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.float64 or not x.is_cuda:
            raise ValueError("device tensors must be CUDA float64")
        return x if x.is_contiguous() else permute(x, list(range(x.dim())))
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    return torch.from_numpy(a).cuda()


def to_host(x) -> np.ndarray:
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def like(result: torch.Tensor, template):
    """Return `result` where `template` lives: numpy in -> numpy out, device in -> device out."""
    return result if isinstance(template, torch.Tensor) else to_host(result)


def empty(shape: Sequence[int], device=None) -> torch.Tensor:
    return torch.empty([int(s) for s in shape], dtype=torch.float64, device=device or "cuda")


def zeros(shape: Sequence[int], device=None) -> torch.Tensor:
    out = empty(shape, device)
    if out.numel():
        check(_lib.lib().ttb_fill_f64(out.data_ptr(), out.numel(), 0.0, _stream_ptr()))
    return out


def strided_op(dst: torch.Tensor, dst_off: int, dst_strides: Sequence[int], src: torch.Tensor, src_off: int,
               src_strides: Sequence[int], shape: Sequence[int], op: int = 0, alpha: float = 1.0) -> None:
    """dst[...] (op)= src[...] over `shape` with explicit element strides and offsets."""
    nd = len(shape)
    if any(int(s) == 0 for s in shape):
        return
    check(
        _lib.lib().ttb_strided_op_f64(
            dst.data_ptr() + 8 * int(dst_off), src.data_ptr() + 8 * int(src_off), nd, _i64(shape),
            _i64(dst_strides), _i64(src_strides), int(op), float(alpha), _stream_ptr(),
        )
    )


def _c_strides(shape: Sequence[int]) -> List[int]:
    st, acc = [], 1
    for n in reversed(list(shape)):
        st.append(acc)
        acc *= int(n)
    return st[::-1]


def permute(x: torch.Tensor, perm: Sequence[int]) -> torch.Tensor:
    """Contiguous copy of x with dimensions reordered: out[i_0..] = x[..i_perm^-1..]
    (np.permute_dims followed by the copy that reshape makes, pytens/algs.py:244-248)."""
    perm = [int(p) for p in perm]
    if sorted(perm) != list(range(x.dim())):
        raise ValueError(f"bad permutation {perm} for a {x.dim()}-d tensor")
    shape = [int(x.shape[p]) for p in perm]
    if perm == list(range(x.dim())) and x.is_contiguous():
        return x
    out = empty(shape, x.device)
    xs = [int(s) for s in x.stride()]
    strided_op(out, 0, _c_strides(shape), x, 0, [xs[p] for p in perm], shape)
    return out


def place(dst: torch.Tensor, src: torch.Tensor, offsets: Sequence[int]) -> None:
    """dst[o_0 : o_0 + s_0, ...] = src (block of a block-diagonal / concatenated core)."""
    ds = [int(s) for s in dst.stride()]
    off = sum(int(o) * s for o, s in zip(offsets, ds))
    strided_op(dst, off, ds, src, 0, [int(s) for s in src.stride()], [int(s) for s in src.shape])


def scal(x: torch.Tensor, alpha: float) -> torch.Tensor:
    """x *= alpha in place (contiguous)."""
    if x.numel():
        check(_lib.lib().ttb_axpby_f64(x.numel(), 0.0, None, float(alpha), x.data_ptr(), _stream_ptr()))
    return x


def axpy(alpha: float, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """y += alpha x in place (both contiguous, same size)."""
    if x.numel() != y.numel():
        raise ValueError("axpy: size mismatch")
    if x.numel():
        check(_lib.lib().ttb_axpby_f64(x.numel(), float(alpha), x.data_ptr(), 1.0, y.data_ptr(), _stream_ptr()))
    return y


def mm(a: torch.Tensor, b: torch.Tensor, ta: bool = False, tb: bool = False) -> torch.Tensor:
    """op(a) @ op(b) for contiguous 2-d tensors on the DMMA GEMM (`ttb_gemm_f64`)."""
    L = _lib.lib()
    m, k = (int(a.shape[1]), int(a.shape[0])) if ta else (int(a.shape[0]), int(a.shape[1]))
    k2, n = (int(b.shape[1]), int(b.shape[0])) if tb else (int(b.shape[0]), int(b.shape[1]))
    if k != k2:
        raise ValueError(f"mm: inner extents differ ({k} vs {k2})")
    out = empty((m, n), a.device)
    if m == 0 or n == 0:
        return out
    if k == 0:
        return zeros((m, n), a.device)
    sam, sak = (1, m) if ta else (k, 1)
    sbk, sbn = (1, k) if tb else (n, 1)
    ws = workspace(L.ttb_gemm_workspace_bytes(m, n, k), a.device, slot="gemm")
    check(
        L.ttb_gemm_f64(
            m, n, k, 1.0, a.data_ptr(), sam, sak, b.data_ptr(), sbk, sbn, 0.0, out.data_ptr(), n,
            ws.data_ptr(), ws.numel(), _stream_ptr(),
        )
    )
    return out


def _matrix_view(x: torch.Tensor, rows: Sequence[int], cols: Sequence[int]) -> Tuple[torch.Tensor, bool]:
    """x (contiguous) as the matrix (prod rows) x (prod cols) without a copy when the dimensions
    already sit in the order rows+cols (plain) or cols+rows (transposed view); otherwise a permuted copy.
    Returns (2-d contiguous tensor, transposed?) with matrix = t.T when transposed."""
    nd = x.dim()
    rows, cols = [int(i) for i in rows], [int(i) for i in cols]
    nr = int(np.prod([x.shape[i] for i in rows], dtype=np.int64)) if rows else 1
    nc = int(np.prod([x.shape[i] for i in cols], dtype=np.int64)) if cols else 1
    if rows + cols == list(range(nd)):
        return x.view(nr, nc), False
    if cols + rows == list(range(nd)):
        return x.view(nc, nr), True
    return permute(x, rows + cols).view(nr, nc), False


def contract(a: torch.Tensor, a_labels: Sequence, b: torch.Tensor, b_labels: Sequence) -> Tuple[torch.Tensor, list]:
    """Contract two tensors over their common labels -- Tensor.contract, pytens/algs.py:201-236
    (a two-operand np.einsum there).  Output order: a's free labels (a's order), then b's free
    labels (b's order).  One GEMM; operands are permuted to matrices only when their layout
    requires it."""
    a_labels, b_labels = list(a_labels), list(b_labels)
    common = [l for l in a_labels if l in b_labels]
    a_free = [i for i, l in enumerate(a_labels) if l not in b_labels]
    b_free = [i for i, l in enumerate(b_labels) if l not in a_labels]
    a_com = [a_labels.index(l) for l in common]
    b_com = [b_labels.index(l) for l in common]
    am, ta = _matrix_view(a, a_free, a_com)  # (free_a x K) or its transpose
    bm, tb = _matrix_view(b, b_com, b_free)  # (K x free_b) or its transpose
    out = mm(am, bm, ta=ta, tb=tb)
    shape = [int(a.shape[i]) for i in a_free] + [int(b.shape[i]) for i in b_free]
    return out.view(shape), [a_labels[i] for i in a_free] + [b_labels[i] for i in b_free]


def qr(mat: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Thin QR of a contiguous (m x n) matrix: q (m x k) with orthonormal columns, r (k x n),
    k = min(m, n) -- np.linalg.qr in Tensor.qr (pytens/algs.py:285).  The factorisation runs in row
    space on the transpose (`ttb_orth_rows_f64`), so q and r are equal to LAPACK's up to the signs of
    the columns of q / rows of r."""
    L = _lib.lib()
    m, n = int(mat.shape[0]), int(mat.shape[1])
    k = min(m, n)
    if m >= n:
        work = permute(mat, [1, 0]) if mat.is_contiguous() else mat.t().contiguous()
        if work.data_ptr() == mat.data_ptr():
            work = work.clone()
        work = work.view(n, m)  # rows of `work` = columns of mat
        R = empty((n, n), mat.device)
        ws = workspace(L.ttb_orth_rows_workspace_bytes(n, m), mat.device)
        check(L.ttb_orth_rows_f64(work.data_ptr(), n, m, R.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
        return permute(work, [1, 0]), R
    # wide: Q R of the leading square block, then R_2 = Q^T M_2 (what Householder QR yields when the
    # leading block has full rank; a rank-deficient leading block takes the orth_rows path on all columns)
    lead = permute(mat[:, :m], [1, 0])  # rows = leading columns of mat
    R1 = empty((m, m), mat.device)
    ws = workspace(L.ttb_orth_rows_workspace_bytes(m, m), mat.device)
    check(L.ttb_orth_rows_f64(lead.data_ptr(), m, m, R1.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
    q = permute(lead, [1, 0])  # m x m
    r = mm(q, mat, ta=True)  # exact R = Q^T M (its leading block equals R1 up to roundoff)
    return q, r


def trunc_svd(mat: torch.Tensor, delta: float, with_normalizing: bool = False):
    """delta-truncated SVD of a contiguous matrix (`ttb_delta_svd_f64`): (u, s, svt, info)."""
    from .utils import delta_svd_dev

    return delta_svd_dev(mat, delta, with_normalizing)


def unscale_rows(svt: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """v = svt / s[:, None] as a new tensor (delta_svd returns v, pytens/utils.py:94-100)."""
    v = svt.clone() if svt.is_contiguous() else permute(svt, [0, 1])
    if v.numel():
        check(_lib.lib().ttb_scale_rows_f64(v.data_ptr(), int(v.shape[0]), int(v.shape[1]), int(v.shape[1]),
                                             s.data_ptr(), 2, _stream_ptr()))
    return v


def diag(s: torch.Tensor) -> torch.Tensor:
    n = int(s.shape[0])
    out = empty((n, n), s.device)
    if n:
        check(_lib.lib().ttb_diag_f64(s.data_ptr(), n, out.data_ptr(), _stream_ptr()))
    return out
