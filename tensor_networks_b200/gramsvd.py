"""Gram-SVD TT rounding on the device -- SURVEY.md section 8(f) row 2.

Mirror of `eps_to_rank` / `gram_eig_and_svd` / `tt_gramsvd_round` of the reference
(pytens/algs.py:1707-1717, :1720-1768, :1771-1838; exercised by tests/main_test.py:245-262).  The
Gram sweeps and all core updates are FP64 GEMMs on the DMMA kernels of this package (`ttb_gemm_f64`);
the symmetric eigendecompositions run on the device Jacobi kernel, batched: all right Gram matrices of a
train are known after the first sweep and are factored by ONE launch (`ttb_gram_eig_batched_f64`, one
thread-block cluster per matrix), which also applies the reference's decimal rounding of the square
roots and writes the scaled eigenvector matrices A = V diag(e12), B = V diag(em12) -- no eigenvalue
visits the host and no elementwise glue runs in eager torch.  Per bond that leaves one eigendecomposition
(the left Gram matrix of the updated core), one truncated SVD (`ttb_delta_svd_f64`) and five GEMMs.
Gram matrices larger than 256 x 256 take the generic path (three `ttb_delta_svd_f64` calls per bond).
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check
from .tt import TensorTrain, _require_cuda, _stream_ptr, workspace
from .utils import delta_svd_dev


def dev_mm(a: torch.Tensor, b: torch.Tensor, ta: bool = False, tb: bool = False) -> torch.Tensor:
    """op(a) @ op(b) for contiguous 2-d CUDA float64 tensors on the package's DMMA GEMM."""
    L = _lib.lib()
    a = a.contiguous()
    b = b.contiguous()
    m, k = (int(a.shape[1]), int(a.shape[0])) if ta else (int(a.shape[0]), int(a.shape[1]))
    k2, n = (int(b.shape[1]), int(b.shape[0])) if tb else (int(b.shape[0]), int(b.shape[1]))
    if k != k2:
        raise ValueError(f"dev_mm: inner extents differ ({k} vs {k2})")
    out = torch.empty((m, n), dtype=torch.float64, device=a.device)
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


def eps_to_rank(s: np.ndarray, eps: float) -> int:
    """Rank of a truncated SVD with tail energy <= eps -- pytens/algs.py:1707-1717."""
    tail = np.sqrt(np.cumsum(np.square(s[::-1])))[::-1] <= eps
    res = int(np.argmax(tail))
    if res == 0 and not tail[0]:
        return int(s.shape[0])
    if res == 0 and tail[0]:
        return 1
    return res


def _rounded_sqrt(eig: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """sqrt(|eig|) rounded at 1e-8 of the largest value and its masked reciprocal
    (pytens/algs.py:1729-1749), returned as device vectors."""
    pos_tol = 1e-15
    e12 = np.sqrt(np.abs(eig.cpu().numpy()))
    threshold = np.ceil(np.log10(np.max(e12) * 1e-8 + pos_tol))
    e12 = np.round(e12, min(-int(threshold), 16))
    em12 = np.zeros_like(e12)
    nz = e12 != 0
    em12[nz] = 1.0 / e12[nz]
    return torch.from_numpy(e12).to(eig.device), torch.from_numpy(em12).to(eig.device)


GRAM_EIG_MAX = 256  # largest Gram matrix of the batched on-chip eigendecomposition


def gram_eig_batched_dev(g: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Eigen-factors of a stack of Gram matrices g (count, p, p), p <= 256, in one launch:
    (A, B, eig) with A = V diag(e12), B = V diag(em12) (pytens/algs.py:1727-1749), eig (count, p) descending."""
    L = _lib.lib()
    g = g.contiguous()
    count, p = int(g.shape[0]), int(g.shape[1])
    a = torch.empty_like(g)
    b = torch.empty_like(g)
    eig = torch.empty((count, p), dtype=torch.float64, device=g.device)
    ws = workspace(L.ttb_gram_eig_batched_workspace_bytes(count, p), g.device, slot="gram_eig")
    check(
        L.ttb_gram_eig_batched_f64(
            g.data_ptr(), count, p, a.data_ptr(), b.data_ptr(), eig.data_ptr(), None, ws.data_ptr(), ws.numel(), _stream_ptr()
        )
    )
    return a, b, eig


def bond_factors_dev(al: torch.Tensor, bl: torch.Tensor, ar: torch.Tensor, br: torch.Tensor, delta: float):
    """(curr, next) of one bond from the eigen-factors of its left / right Gram matrices (pytens/algs.py:1751-1768):
    tmp = A_l^T A_r, (u, s v^T) = truncated SVD, curr = B_l u, next = s v^T B_r^T."""
    tmp = dev_mm(al, ar, ta=True)
    u, _s, svt, _ = delta_svd_dev(tmp, float(delta))  # rank rule == eps_to_rank (tail energy <= delta)
    return dev_mm(bl, u), dev_mm(svt, br, tb=True)


def gram_eig_and_svd_dev(gl: torch.Tensor, gr: torch.Tensor, delta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Device form of gram_eig_and_svd (pytens/algs.py:1720-1768): returns (curr, next) with
    curr (r x rk) to be applied to core i from the right and next (rk x r) to core i+1 from the left."""
    if max(int(gl.shape[0]), int(gr.shape[0])) <= GRAM_EIG_MAX:
        al, bl, _ = gram_eig_batched_dev(gl[None])
        ar, br, _ = gram_eig_batched_dev(gr[None])
        return bond_factors_dev(al[0], bl[0], ar[0], br[0], delta)
    vl, eigl, _, _ = delta_svd_dev(gl, 0.0)  # G = V |Lambda| V^T: columns of vl are the eigenvectors
    vr, eigr, _, _ = delta_svd_dev(gr, 0.0)
    eigl12, eiglm12 = _rounded_sqrt(eigl)
    eigr12, eigrm12 = _rounded_sqrt(eigr)
    tmp = dev_mm(vl * eigl12[None, :], vr * eigr12[None, :], ta=True)  # (L^1/2 Vl^T)(Vr R^1/2)
    u, _s, svt, _ = delta_svd_dev(tmp, float(delta))  # rank rule == eps_to_rank (tail energy <= delta)
    curr = dev_mm(vl, eiglm12[:, None] * u)
    nxt = dev_mm(svt * eigrm12[None, :], vr, tb=True)
    return curr, nxt


def gram_eig_and_svd(gl: np.ndarray, gr: np.ndarray, delta: float) -> Tuple[np.ndarray, np.ndarray]:
    """Drop-in for `pytens.algs.gram_eig_and_svd` (host arrays in and out, computed on the GPU)."""
    _require_cuda()
    a = torch.from_numpy(np.ascontiguousarray(gl, dtype=np.float64)).cuda()
    b = torch.from_numpy(np.ascontiguousarray(gr, dtype=np.float64)).cuda()
    c, n = gram_eig_and_svd_dev(a, b, delta)
    return c.cpu().numpy(), n.cpu().numpy()


def gramsvd_round(tt: TensorTrain, eps: float) -> TensorTrain:
    """Round `tt` in place by Gram SVD -- tt_gramsvd_round, pytens/algs.py:1771-1838.

    Right Gram matrices by a right-to-left sweep (:1808-1815), delta = eps ||X|| / sqrt(d-1) from the
    last of them (:1817-1818), then per bond: left Gram of the already updated core, the Gram
    eigen/SVD step, and the two core updates (:1822-1836).  `tt.last_gramsvd` holds {"delta", "norm"}."""
    _require_cuda()
    d = tt.d
    if d < 2:
        raise ValueError("gramsvd_round needs at least two cores")
    cores = tt.cores
    last = cores[d - 1].reshape(cores[d - 1].shape[0], -1)
    gr = [None] * d
    gr[d - 1] = dev_mm(last, last, tb=True)
    for i in range(d - 2, -1, -1):
        c = cores[i]
        r0, n, r1 = (int(x) for x in c.shape)
        tmp = dev_mm(c.reshape(r0 * n, r1), gr[i + 1]).reshape(r0, n * r1)
        gr[i] = dev_mm(tmp, c.reshape(r0, n * r1), tb=True)
    norm = float(np.sqrt(gr[0].reshape(-1)[0].item()))
    delta = eps * norm / (d - 1) ** 0.5
    # all right Gram matrices at once, one launch per distinct size (usually one)
    right = {}
    by_size = {}
    for i in range(1, d):
        by_size.setdefault(int(gr[i].shape[0]), []).append(i)
    for size, idxs in by_size.items():
        if size <= GRAM_EIG_MAX:
            a, b, _ = gram_eig_batched_dev(torch.stack([gr[i] for i in idxs]))
            for pos, i in enumerate(idxs):
                right[i] = (a[pos], b[pos])
    for i in range(d - 1):
        c = cores[i]
        r0, n, r1 = (int(x) for x in c.shape)
        m2 = c.reshape(r0 * n, r1)
        gl = dev_mm(m2, m2, ta=True)
        if (i + 1) in right and r1 <= GRAM_EIG_MAX:
            al, bl, _ = gram_eig_batched_dev(gl[None])
            curr, nxt = bond_factors_dev(al[0], bl[0], right[i + 1][0], right[i + 1][1], delta)
        else:
            curr, nxt = gram_eig_and_svd_dev(gl, gr[i + 1], delta)
        rk = int(curr.shape[1])
        cores[i] = dev_mm(m2, curr).reshape(r0, n, rk)
        c1 = cores[i + 1]
        cores[i + 1] = dev_mm(nxt, c1.reshape(int(c1.shape[0]), -1)).reshape(rk, int(c1.shape[1]), int(c1.shape[2]))
    tt.cores = cores
    tt.last_gramsvd = {"delta": float(delta), "norm": norm}
    return tt
