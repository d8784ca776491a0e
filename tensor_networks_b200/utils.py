"""delta-truncated SVD on the device -- mirror of `pytens/utils.py`."""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import check
from .tt import _require_cuda, _stream_ptr, workspace


@dataclass
class TruncSVD:
    """Store a truncated SVD (pytens/utils.py:8-16)."""

    u: np.ndarray
    s: np.ndarray
    v: np.ndarray
    remaining_delta: float
    delta: Optional[float] = None


def delta_svd_dev(data: torch.Tensor, delta: float, with_normalizing: bool = False, max_rank: int = 0):
    """Device form: returns (u, s, svt, info) as CUDA tensors; svt = diag(s) @ v."""
    _require_cuda()
    L = _lib.lib()
    if data.dim() != 2 or data.dtype != torch.float64 or not data.is_cuda:
        raise ValueError("delta_svd_dev needs a 2-d CUDA float64 tensor")
    data = data.contiguous()
    m, n = int(data.shape[0]), int(data.shape[1])
    p = min(m, n)
    u = torch.empty(m * p, dtype=torch.float64, device=data.device)
    s = torch.empty(p, dtype=torch.float64, device=data.device)
    svt = torch.empty(p * n, dtype=torch.float64, device=data.device)
    info = (ctypes.c_double * 4)()
    ws = workspace(L.ttb_delta_svd_workspace_bytes(m, n), data.device)
    check(
        L.ttb_delta_svd_f64(
            data.data_ptr(), m, n, float(delta), 1 if with_normalizing else 0, int(max_rank), u.data_ptr(),
            s.data_ptr(), svt.data_ptr(), info, ws.data_ptr(), ws.numel(), _stream_ptr(),
        )
    )
    rank = int(info[0])
    return (
        u[: m * rank].view(m, rank),
        s[:rank],
        svt[: rank * n].view(rank, n),
        {"rank": rank, "delta": float(info[1]), "remaining_delta": float(info[2]), "fro2": float(info[3])},
    )


def delta_svd(data: np.ndarray, delta: float, with_normalizing: bool = False) -> TruncSVD:
    """Drop-in for `pytens.utils.delta_svd` (pytens/utils.py:19-100), computed on the GPU.

    Same truncation rule and return type; u/v are determined up to the usual sign
    (and, for repeated singular values, rotation) ambiguity of an SVD.
    """
    dev = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64)).cuda()
    u, s, svt, info = delta_svd_dev(dev, delta, with_normalizing)
    s_h = s.cpu().numpy()
    svt_h = svt.cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.where(s_h[:, None] > 0, svt_h / s_h[:, None], 0.0)
    return TruncSVD(
        u.cpu().numpy(),
        s_h,
        v,
        info["remaining_delta"],
        info["delta"] if with_normalizing else None,
    )


def orth_rows_dev(mat: torch.Tensor):
    """Orthonormalise the rows of `mat` (c x m CUDA float64) in place; returns (Q, R) with
    mat_in^T = Q^T R -- the device form of np.linalg.qr(mat.T) in tt_right_orth
    (pytens/algs.py:1678) without the transposed copy."""
    _require_cuda()
    L = _lib.lib()
    if mat.dim() != 2 or mat.dtype != torch.float64 or not mat.is_cuda or not mat.is_contiguous():
        raise ValueError("orth_rows_dev needs a contiguous 2-d CUDA float64 tensor")
    c, m = int(mat.shape[0]), int(mat.shape[1])
    R = torch.empty((c, c), dtype=torch.float64, device=mat.device)
    ws = workspace(L.ttb_orth_rows_workspace_bytes(c, m), mat.device)
    check(L.ttb_orth_rows_f64(mat.data_ptr(), c, m, R.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return mat, R
