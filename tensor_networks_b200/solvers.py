"""TT operators and TT-GMRES on the device -- the heaviest in-repo caller of the hot path.

Mirror of `ttop_rank1` / `ttop_apply` / `gmres` of the reference (pytens/algs.py:2383-2420,
:2662-2697, :2701-2793; exercised by tests/main_test.py:428-448).  Everything stays resident in HBM:
the Arnoldi loop calls `TensorTrain.round` (tt_svd_round), `inner` and `norm` -- the kernels of this
package -- per step, and the operator application is one DMMA GEMM per core between permuted operands.  Only the small Hessenberg least-squares problem runs on the host, as in the
reference.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import dense
from .tt import TensorTrain, _require_cuda


class TTOperator:
    """A TT-matrix: cores (r_{k-1}, n_out_k, n_in_k, r_k) with r_0 = r_d = 1, on the device.

    Same index order as the reference's interior operator cores (`pytens/algs.py:2398-2407`:
    rank, out, in, rank)."""

    def __init__(self, cores: Sequence[torch.Tensor]):
        _require_cuda()
        cores = [c.contiguous() for c in cores]
        for k, c in enumerate(cores):
            if c.dim() != 4 or c.dtype != torch.float64 or not c.is_cuda:
                raise ValueError(f"operator core {k}: need a 4-d CUDA float64 tensor")
        if cores[0].shape[0] != 1 or cores[-1].shape[3] != 1:
            raise AssertionError("boundary operator ranks must be 1")
        for k in range(len(cores) - 1):
            if cores[k].shape[3] != cores[k + 1].shape[0]:
                raise AssertionError(f"operator bond {k} does not chain")
        self.cores: List[torch.Tensor] = cores

    @property
    def d(self) -> int:
        return len(self.cores)

    def dense(self) -> np.ndarray:
        """Dense (out_1, in_1, ..., out_d, in_d) array (small operators only; test helper)."""
        acc = self.cores[0]
        for c in self.cores[1:]:
            acc = torch.tensordot(acc, c, dims=([acc.dim() - 1], [0]))
        return acc.squeeze(0).squeeze(-1).cpu().numpy()


def ttop_rank1(mats: Sequence[np.ndarray], device="cuda") -> TTOperator:
    """Rank-1 TT operator A_1 (x) A_2 (x) ... (x) A_d -- `ttop_rank1`, pytens/algs.py:2383-2434."""
    cores = []
    for m in mats:
        a = torch.from_numpy(np.ascontiguousarray(np.asarray(m, dtype=np.float64))).to(device)
        cores.append(a.reshape(1, a.shape[0], a.shape[1], 1))
    return TTOperator(cores)


def ttop_apply(op: TTOperator, tt: TensorTrain) -> TensorTrain:
    """y = op(x): core_k[(m, i), j, (p, l)] = sum_k op_k[i, j, k, l] x_k[m, k, p].

    Bond ordering (TT rank major, operator rank minor) as in `ttop_apply`,
    pytens/algs.py:2662-2697 ("ijkl,mkp->mijpl" then reshape)."""
    if op.d != tt.d:
        raise AssertionError("operator and tensor train have different lengths")
    out = []
    for a, v in zip(op.cores, tt.cores):
        if a.shape[2] != v.shape[1]:
            raise AssertionError("operator input mode size does not match the tensor train")
        c, _ = dense.contract(a, "ijkl", v, "mkp")  # one DMMA GEMM -> (i, j, l, m, p)
        c = dense.permute(c, [3, 0, 1, 4, 2])  # (m, i, j, p, l)
        s = [int(x) for x in c.shape]
        out.append(c.reshape(s[0] * s[1], s[2], s[3] * s[4]))
    return TensorTrain(out)


def gmres(
    op: Callable[[TensorTrain], TensorTrain],
    rhs: TensorTrain,
    x0: TensorTrain,
    eps: float = 1e-5,
    round_eps: float = 1e-10,
    maxiter: int = 100,
) -> Tuple[TensorTrain, float]:
    """TT-GMRES with rounding after every operator application and orthogonalisation step.

    Follows `gmres` of the reference line by line (pytens/algs.py:2701-2793): modified Gram-Schmidt
    against the Krylov basis with `inner`, `tt_svd_round(w, round_eps)` twice per step, `norm`, a dense
    least-squares solve of the Hessenberg system on the host, stop when the least-squares residual is
    below eps.  Returns (x, ||rhs - op(x)||)."""
    r0 = (rhs + op(x0).scale(-1.0)).round(round_eps).compact()  # re-own: do not pin the unrounded buffers
    beta = float(r0.norm())
    if beta == 0.0:
        return x0.clone(), 0.0
    v = [r0.scale(1.0 / beta)]
    y: Optional[np.ndarray] = None
    H = np.zeros((0, 0))
    for jj in range(maxiter):
        w = op(v[-1]).round(round_eps)
        newH = np.zeros((jj + 2, jj + 1))
        newH[: H.shape[0], : H.shape[1]] = H
        H = newH
        for ii in range(jj + 1):
            H[ii, jj] = float(w.inner(v[ii]))
            w = w + v[ii].clone().scale(-H[ii, jj])
        w = w.round(round_eps).compact()
        H[jj + 1, jj] = float(w.norm())
        if H[jj + 1, jj] > 0.0:
            v.append(w.scale(1.0 / H[jj + 1, jj]))
        e = np.zeros(H.shape[0])
        e[0] = beta
        y, resid, _, _ = np.linalg.lstsq(H, e, rcond=None)
        if H[jj + 1, jj] == 0.0 or (resid.size > 0 and abs(float(resid[0])) < eps):
            break
    x = x0.clone()
    for vv, yy in zip(v, y):
        x = x + vv.clone().scale(float(yy))
    x = x.round(round_eps)
    resid_tt = rhs + op(x).scale(-1.0)
    return x, float(resid_tt.norm())
