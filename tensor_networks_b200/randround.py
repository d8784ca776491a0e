"""Randomised TT rounding on the device -- SURVEY.md section 8(f) row 4.

Mirror of `TTRandRound` / `tt_randomized_round` / `tt_sum_randomized_round` /
`tt_rand_precond_svd_round` (pytens/algs.py:2133-2380; Daas et al., arXiv:2110.04393, algorithms 3.2
and 3.4).  The Gaussian sketch cores are drawn on the host with the same `np.random.randn` calls, in
the same order and with the same normalisation as the reference (so a seeded run sketches with the
same matrices); everything after that -- the right-to-left partial contractions, the sketch products,
the QR factorisations and the projections -- runs on the device: DMMA GEMMs (`ttb_gemm_f64`) and
the row-space QR (`ttb_orth_rows_f64`).  The rounded TT spans the same subspaces as the reference's
(a QR basis is unique up to signs), so the represented tensor agrees to roundoff.
"""

from __future__ import annotations

import copy
from typing import List, Optional, Union

import numpy as np

from . import dense


def _algs():
    from . import algs

    return algs


def _mat(t, rows: int):
    return t.reshape(rows, -1)


class TTRandRound:
    """Randomised rounding of a TT (`y` a TensorNetwork) or of a sum of TTs (`y` a list) to
    `target_ranks` -- pytens/algs.py:2133-2314."""

    def __init__(self, y, target_ranks: List):
        A = _algs()
        self.y = y
        self.target_ranks = target_ranks
        if isinstance(y, list) and len(y) > 0 and hasattr(y[0], "network"):
            self.ns = len(y)
            self.d = y[0].network.number_of_nodes()
        elif hasattr(y, "network"):
            self.ns = 1
            self.d = y.network.number_of_nodes()
        else:
            raise ValueError(f"Invalid type for y ({type(y)}). Argument y only accepts a list of "
                             "TensorNetworks or a TensorNetwork")
        del A

    def init_rand_mat(self, ranks: Optional[List] = None) -> List[np.ndarray]:
        """Gaussian TT cores scaled by 1/sqrt(size) -- pytens/algs.py:2164-2182 (host draw, same order)."""
        if ranks is None:
            ranks = self.target_ranks
        sh = self.y[0].shape() if isinstance(self.y, list) else self.y.shape()
        cores = []
        for i in range(self.d):
            if i == 0:
                shp = [sh[i], ranks[i]]
            elif i == self.d - 1:
                shp = [ranks[i - 1], sh[i]]
            else:
                shp = [ranks[i - 1], sh[i], ranks[i]]
            cores.append(np.random.randn(*shp) / np.sqrt(np.prod(shp)))
        return cores

    def partial_contraction(self, tt, y: list, direction: str = "rl") -> list:
        """w_i = contraction of cores i.. of `tt` with cores i.. of the sketch `y` (r_i x l_i),
        right to left -- pytens/algs.py:2184-2209.  Two GEMMs per core."""
        if direction != "rl":
            raise ValueError("Invalid option")
        host = not dense.is_dev(tt.value(self.d - 1))
        w = []
        for i in range(self.d - 1, 0, -1):
            x = dense.as_dev(tt.value(i))
            yi = dense.as_dev(y[i])
            if i == self.d - 1:
                w.append(dense.mm(x, yi, tb=True))  # (r x n)(l x n)^T
                continue
            r0, r1 = int(x.shape[0]), int(x.shape[-1])
            tmp = dense.mm(x.reshape(-1, r1), w[-1]).reshape(r0, -1)  # (r0, n l')
            w.append(dense.mm(tmp, yi.reshape(int(yi.shape[0]), -1), tb=True))  # (r0, l)
        w = w[::-1]
        return [dense.to_host(m) for m in w] if host else w

    def rand_then_orth(self):
        """Algorithm 3.2 (randomise-then-orthogonalise) -- pytens/algs.py:2211-2238."""
        if isinstance(self.y, list):
            raise ValueError("It seems that this function is being used to round a TT-sum")
        y = self.y
        host = not dense.is_dev(y.value(0))
        sketch = self.init_rand_mat()
        w = [dense.as_dev(m) for m in self.partial_contraction(y, sketch, "rl")]
        res = copy.deepcopy(y)
        x = dense.as_dev(y.value(0))
        for i in range(self.d - 1):
            lead = [int(s) for s in x.shape[:-1]]
            zn = x.reshape(-1, int(x.shape[-1]))
            q, _ = dense.qr(dense.mm(zn, w[i]))
            core = q.reshape(lead + [int(q.shape[-1])])
            res.network.nodes[i]["tensor"].update_val_size(dense.to_host(core) if host else core)
            nxt = dense.as_dev(y.value(i + 1))
            tail = [int(s) for s in nxt.shape[1:]]
            proj = dense.mm(q, zn, ta=True)  # q^T zn
            x = dense.mm(proj, nxt.reshape(int(nxt.shape[0]), -1)).reshape([int(q.shape[-1])] + tail)
        res.network.nodes[self.d - 1]["tensor"].update_val_size(dense.to_host(x) if host else x)
        return res

    def rto_rounding_ttsum(self):
        """Algorithm 3.4: randomise-then-orthogonalise for a sum of TTs without forming the sum --
        pytens/algs.py:2240-2306."""
        if not isinstance(self.y, list):
            raise ValueError("It seems that this function is being used to round a single TT")
        ys = self.y
        host = not dense.is_dev(ys[0].value(0))
        sketch = self.init_rand_mat()
        w = [[dense.as_dev(m) for m in self.partial_contraction(t, sketch)] for t in ys]
        res = copy.deepcopy(ys[0])
        firsts = [dense.as_dev(t.value(0)) for t in ys]
        n0 = int(firsts[0].shape[0])
        offs = np.cumsum([0] + [int(f.shape[1]) for f in firsts])
        x = dense.empty((n0, int(offs[-1])))
        for j, f in enumerate(firsts):
            dense.place(x, f, (0, int(offs[j])))
        for i in range(self.d - 1):
            lead = [int(s) for s in x.shape[:-1]]
            rk = [int(t.value(i).shape[-1]) for t in ys]
            rkp1 = [int(t.value(i + 1).shape[-1]) for t in ys]
            cum = np.cumsum([0] + rk)
            wcat = dense.empty((int(cum[-1]), int(w[0][i].shape[1])))
            for j in range(self.ns):
                dense.place(wcat, w[j][i], (int(cum[j]), 0))
            zn = x.reshape(-1, int(x.shape[-1]))
            q, _ = dense.qr(dense.mm(zn, wcat))
            self.target_ranks[i] = min(self.target_ranks[i], int(q.shape[-1]))
            rho = self.target_ranks[i]
            mn = dense.mm(q, zn, ta=True)
            core = q.reshape(lead + [rho])
            res.network.nodes[i]["tensor"].update_val_size(dense.to_host(core) if host else core)
            n_next = int(ys[0].value(i + 1).shape[1])
            if i < self.d - 2:
                tot = int(sum(rkp1))
                x = dense.empty((rho, n_next, tot))
                co = np.cumsum([0] + rkp1)
                for j in range(self.ns):
                    nxt = dense.as_dev(ys[j].value(i + 1))
                    blk = dense.permute(mn[:, int(cum[j]):int(cum[j + 1])], [0, 1])
                    part = dense.mm(blk, nxt.reshape(int(nxt.shape[0]), -1)).reshape(rho, n_next, rkp1[j])
                    dense.place(x, part, (0, 0, int(co[j])))
            else:
                x = dense.zeros((rho, n_next))
                for j in range(self.ns):
                    nxt = dense.as_dev(ys[j].value(i + 1))
                    blk = dense.permute(mn[:, int(cum[j]):int(cum[j + 1])], [0, 1])
                    dense.axpy(1.0, dense.mm(blk, nxt.reshape(int(nxt.shape[0]), -1)), x)
                res.network.nodes[self.d - 1]["tensor"].update_val_size(dense.to_host(x) if host else x)
        return res

    def round(self):
        return self.rto_rounding_ttsum() if isinstance(self.y, list) else self.rand_then_orth()


def tt_randomized_round(y, target_ranks: List):
    """pytens/algs.py:2317-2321."""
    return TTRandRound(y, target_ranks).rand_then_orth()


def tt_sum_randomized_round(y: list, target_ranks: List):
    """pytens/algs.py:2324-2330."""
    return TTRandRound(y, target_ranks).rto_rounding_ttsum()


def tt_rand_precond_svd_round(tn, eps: float, rank_bound: List[int]):
    """Randomised rounding to `rank_bound` followed by a right-to-left delta-truncated SVD sweep with
    delta = eps / sqrt(d-1) relative to each unfolding -- pytens/algs.py:2333-2380."""
    setup = TTRandRound(y=tn, target_ranks=rank_bound)
    res = setup.round()
    dim = setup.d
    host = not dense.is_dev(res.value(0))
    for i in range(dim - 1, 0, -1):
        cur = dense.as_dev(res.value(i))
        sh = [int(s) for s in cur.shape]
        prev = dense.as_dev(res.value(i - 1))
        delta = eps / (dim - 1) ** 0.5
        u, s, svt, _ = dense.trunc_svd(cur.reshape(sh[0], -1), delta, True)
        v = dense.unscale_rows(svt, s)
        new_cur = v.reshape([int(v.shape[0])] + sh[1:])
        us = dense.mm(u, dense.diag(s))
        ps = [int(x) for x in prev.shape]
        new_prev = dense.mm(prev.reshape(-1, ps[-1]), us).reshape(ps[:-1] + [int(us.shape[1])])
        res.network.nodes[i]["tensor"].update_val_size(dense.to_host(new_cur) if host else new_cur)
        res.network.nodes[i - 1]["tensor"].update_val_size(dense.to_host(new_prev) if host else new_prev)
    return res
