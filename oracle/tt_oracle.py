"""CPU oracle for the tensor-train core-sweep hot path -- TEST INFRASTRUCTURE ONLY.

A numpy (fp64) restatement of the reference's algorithms for the path named in
BASELINE.json (TT inner product / norm, TT rounding, TT-SVD).  It is the
checker for the CUDA path: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The
product package (`tensor_networks_b200`) never does.

Parity status: PINNED.  `tests/test_oracle.py` checks every function here
against golden vectors produced by the reference itself (imported through
`oracle/refshim.py` in the build container, script `oracle/make_golden.py`,
fixtures under `tests/golden/`), and -- when /root/reference is present --
against the live reference side by side.

A tensor train is a list of d numpy arrays ("cores"), core k of shape
(r_{k-1}, n_k, r_k) with r_0 = r_d = 1, C-order -- byte-identical to the
reference's cores (first core (n_1, r_1), last core (r_{d-1}, n_d);
pytens/algs.py:1188-1216) viewed with the unit bonds made explicit.

Each function cites the reference lines it follows.
"""

from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

Cores = List[np.ndarray]


# ----------------------------------------------------------------------------
# construction helpers
# ----------------------------------------------------------------------------
def as_cores3(cores: Sequence[np.ndarray]) -> Cores:
    """View reference-shaped cores (2-D first/last) as (r_{k-1}, n_k, r_k)."""
    d = len(cores)
    out = []
    for k, c in enumerate(cores):
        c = np.asarray(c, dtype=np.float64)
        if c.ndim == 3:
            out.append(c)
        elif c.ndim == 2 and k == 0 and d > 1:
            out.append(c.reshape(1, c.shape[0], c.shape[1]))
        elif c.ndim == 2 and k == d - 1:
            out.append(c.reshape(c.shape[0], c.shape[1], 1))
        else:
            raise ValueError(f"core {k} has unsupported shape {c.shape}")
    return out


def rand_tt(
    shape: Sequence[int],
    ranks: Sequence[int],
    rng: np.random.Generator,
    scaled: bool = True,
) -> Cores:
    """Random TT with standard-normal cores (layout of pytens/algs.py:1180-1218).

    `scaled=True` multiplies core k by (n_k * r_k)^(-1/2) so that ||X|| = O(1)
    (SURVEY.md section 8d); `scaled=False` is the reference-native unscaled variant.
    """
    d = len(shape)
    assert len(ranks) == d - 1
    r = [1] + [int(x) for x in ranks] + [1]
    cores = []
    for k in range(d):
        c = rng.standard_normal((r[k], int(shape[k]), r[k + 1]))
        if scaled:
            c *= 1.0 / math.sqrt(shape[k] * r[k + 1])
        cores.append(c)
    return cores


def tt_add(x: Cores, y: Cores) -> Cores:
    """Formal TT sum X + Y by block-diagonal rank growth.

    Follows TensorNetwork.__add__ -> Tensor.block_diagonal
    (pytens/algs.py:1339-1353, :308-344): first core concatenated along the
    right bond, last core along the left bond, interior cores block-diagonal.
    """
    d = len(x)
    assert d == len(y)
    out = []
    for k in range(d):
        a, b = x[k], y[k]
        assert a.shape[1] == b.shape[1]
        if d == 1:
            out.append(a + b)
        elif k == 0:
            out.append(np.concatenate([a, b], axis=2))
        elif k == d - 1:
            out.append(np.concatenate([a, b], axis=0))
        else:
            c = np.zeros(
                (a.shape[0] + b.shape[0], a.shape[1], a.shape[2] + b.shape[2])
            )
            c[: a.shape[0], :, : a.shape[2]] = a
            c[a.shape[0] :, :, a.shape[2] :] = b
            out.append(c)
    return out


def to_dense(cores: Cores) -> np.ndarray:
    """Contract the chain into the dense tensor (what `contract().value` gives
    for a TT, pytens/algs.py:469-485)."""
    acc = cores[0].reshape(-1, cores[0].shape[2])
    for c in cores[1:]:
        acc = acc @ c.reshape(c.shape[0], -1)
        acc = acc.reshape(-1, c.shape[2])
    return acc.reshape([c.shape[1] for c in cores])


def ranks_of(cores: Cores) -> List[int]:
    return [c.shape[2] for c in cores[:-1]]


# ----------------------------------------------------------------------------
# inner product / norm
# ----------------------------------------------------------------------------
def inner(a: Cores, b: Cores) -> np.ndarray:
    """<A, B> by the left-to-right environment sweep.

    Semantics of TensorNetwork.inner (pytens/algs.py:585-587): attach() shares
    the free indices of the two networks (:521-572) and contract() sums them
    out (:469-485).  For two TTs this is
        E_1 = A_1^T B_1,  E_k = sum_n A_k[:, n, :]^T E_{k-1} B_k[:, n, :],
    with the result E_d (1 x 1).  Returns a 0-d float64 array like the
    reference.
    """
    assert len(a) == len(b)
    env = np.ones((1, 1))
    for ca, cb in zip(a, b):
        ra, n, ra2 = ca.shape
        rb, nb, rb2 = cb.shape
        assert n == nb
        t = env @ cb.reshape(rb, n * rb2)  # (ra, n*rb2)
        env = ca.reshape(ra * n, ra2).T @ t.reshape(ra * n, rb2)  # (ra2, rb2)
    return np.asarray(env[0, 0])


def norm(a: Cores) -> float:
    """sqrt(|<A, A>|) -- TensorNetwork.norm, pytens/algs.py:589-594."""
    return float(np.sqrt(np.abs(float(inner(a, a)))))


# ----------------------------------------------------------------------------
# delta-truncated SVD
# ----------------------------------------------------------------------------
def delta_svd(
    data: np.ndarray, delta: float, with_normalizing: bool = False
) -> Tuple[np.ndarray, np.ndarray, np.ndarray, float, Optional[float]]:
    """Thin SVD + tail-energy truncation; pytens/utils.py:19-100.

    Tall-skinny (m > 10 n) goes through QR first (:56-60), otherwise a direct
    thin SVD (:63).  With `with_normalizing`, delta is scaled by ||data||_F
    computed from the singular values (:70-72).  Trailing singular values are
    dropped while their cumulative energy stays `<= delta**2` (:74-82); at
    least rank 1 is kept (:84).  Returns (u, s, v, remaining_delta, delta_or_None).
    """
    m, n = data.shape
    if m > 10 * n:
        q, r = np.linalg.qr(data)
        u, s, v = np.linalg.svd(r)
        u = q @ u
    else:
        u, s, v = np.linalg.svd(data, full_matrices=False)
    if with_normalizing:
        delta = delta * float(np.sqrt(np.sum(s**2)))
    tail = np.cumsum((s * s)[::-1])
    ndrop = 0
    for val in tail:
        if val <= delta**2:
            ndrop += 1
        else:
            break
    rank = max(len(s) - ndrop, 1)
    used = float(tail[ndrop - 1]) if ndrop > 0 else 0.0
    rem = float(np.sqrt(delta**2 - used))
    return (
        u[:, :rank],
        s[:rank],
        v[:rank, :],
        rem,
        (delta if with_normalizing else None),
    )


# ----------------------------------------------------------------------------
# rounding
# ----------------------------------------------------------------------------
def right_orth(cores: Cores, k: int) -> Cores:
    """One RQ step on core k, in place; pytens/algs.py:1654-1704.

    QR of the transposed horizontal unfolding (:1674-1678); if n*b < r the
    factors are zero-padded so the bond rank is NOT reduced (:1679-1685); the
    last core instead shrinks to min(r, n) (:1695-1697).  R^T is pushed into
    core k-1 (:1699-1702).
    """
    d = len(cores)
    c = cores[k]
    r, n, b = c.shape
    mat = c.reshape(r, n * b)
    q, rr = np.linalg.qr(mat.T, mode="reduced")
    if k < d - 1 or d == 1:
        if q.shape[1] < r:
            q2 = np.zeros((q.shape[0], r))
            q2[:, : q.shape[1]] = q
            r2 = np.zeros((r, rr.shape[1]))
            r2[: rr.shape[0], :] = rr
            q, rr = q2, r2
        cores[k] = q.T.reshape(r, n, b)
    else:
        cores[k] = q.T.reshape(q.shape[1], n, b)
    prev = cores[k - 1]
    cores[k - 1] = np.dot(prev, rr.T)
    return cores


def svd_round(cores: Cores, eps: float) -> Tuple[Cores, float]:
    """TT rounding: RQ pass then left-to-right delta-truncated SVD sweep.

    Follows tt_svd_round, pytens/algs.py:1841-1903.  `eps` is relative:
    delta = eps / sqrt(d-1) * ||X||_F, taken from the singular values of the
    first core after the RQ pass (:1874-1875) and reused for every later core
    (:1893).  The last core only receives the carry (:1889).  Mutates and
    returns `cores`, plus the absolute delta used.
    """
    d = len(cores)
    assert d >= 2
    right_orth(cores, d - 1)
    for j in range(d - 2, 0, -1):
        right_orth(cores, j)

    c0 = cores[0]
    mat = c0.reshape(c0.shape[0] * c0.shape[1], c0.shape[2])
    u, s, v, _rem, delta = delta_svd(mat, eps / np.sqrt(d - 1), with_normalizing=True)
    assert delta is not None
    carry = np.dot(np.diag(s), v)
    cores[0] = u.reshape(c0.shape[0], c0.shape[1], u.shape[1])
    nxt = cores[1]
    cores[1] = np.einsum("ij,jk...->ik...", carry, nxt)

    for k in range(1, d - 1):
        c = cores[k]
        r1, n, r2a = c.shape
        u, s, v, _rem, _ = delta_svd(c.reshape(r1 * n, r2a), delta)
        carry = np.dot(np.diag(s), v)
        cores[k] = u.reshape(r1, n, u.shape[1])
        cores[k + 1] = np.einsum("ij,jk...->ik...", carry, cores[k + 1])
    return cores, float(delta)


# ----------------------------------------------------------------------------
# Gram-SVD rounding
# ----------------------------------------------------------------------------
def eps_to_rank(s: np.ndarray, eps: float) -> int:
    """Rank kept by a truncated SVD with tail energy <= eps (eps_to_rank, pytens/algs.py:1707-1717)."""
    tail = np.sqrt(np.cumsum(np.square(s[::-1])))[::-1] <= eps
    res = int(np.argmax(tail))
    if res == 0 and not tail[0]:
        return int(s.shape[0])
    if res == 0 and tail[0]:
        return 1
    return res


def round_sqrt_eigs(eig: np.ndarray) -> np.ndarray:
    """sqrt(|eig|) rounded to the decimal position of 1e-8 of the largest value
    (pytens/algs.py:1729-1739): square roots of noise-level eigenvalues become exactly 0."""
    pos_tol = 1e-15
    e12 = np.sqrt(np.abs(eig))
    threshold = np.ceil(np.log10(np.max(e12) * 1e-8 + pos_tol))
    return np.round(e12, min(-int(threshold), 16))


def gram_eig_and_svd(gl: np.ndarray, gr: np.ndarray, delta: float) -> Tuple[np.ndarray, np.ndarray]:
    """Low-rank factors of one bond from its left / right Gram matrices
    (gram_eig_and_svd, pytens/algs.py:1720-1768)."""
    eigl, vl = np.linalg.eigh(gl)
    eigr, vr = np.linalg.eigh(gr)
    eigl12 = round_sqrt_eigs(eigl)
    eigr12 = round_sqrt_eigs(eigr)
    eiglm12 = np.zeros_like(eigl12)
    eigrm12 = np.zeros_like(eigr12)
    eiglm12[eigl12 != 0] = 1.0 / eigl12[eigl12 != 0]
    eigrm12[eigr12 != 0] = 1.0 / eigr12[eigr12 != 0]
    tmp = (eigl12[:, None] * vl.T) @ (vr * eigr12[None, :])
    u, s, v = np.linalg.svd(tmp)
    rk = min(tmp.shape[0], tmp.shape[1], eps_to_rank(s, delta))
    curr = vl @ (eiglm12[:, None] * u[:, :rk])
    nxt = (s[:rk, None] * v[:rk] * eigrm12[None, :]) @ vr.T
    return curr, nxt


def gramsvd_round(cores: Cores, eps: float) -> Tuple[Cores, float]:
    """Gram-SVD TT rounding (tt_gramsvd_round, pytens/algs.py:1771-1838): right Gram matrices by a
    right-to-left sweep (:1808-1815), delta = eps ||X|| / sqrt(d-1) from the last of them (:1817-1818),
    then per bond the left Gram of the updated core, gram_eig_and_svd, and the two core updates
    (:1822-1836).  Mutates and returns `cores` (3-d form) plus the absolute delta."""
    d = len(cores)
    assert d >= 2
    last = cores[d - 1].reshape(cores[d - 1].shape[0], -1)
    gr = [last @ last.T]
    for i in range(d - 2, -1, -1):
        c = cores[i]
        r0, n, r1 = c.shape
        tmp = (c.reshape(-1, r1) @ gr[-1]).reshape(r0, n * r1)
        gr.append(tmp @ c.reshape(r0, n * r1).T)
    norm = np.sqrt(gr[-1])[0, 0]
    delta = eps * norm / (d - 1) ** 0.5
    gr = gr[::-1]
    for i in range(d - 1):
        c = cores[i]
        r0, n, r1 = c.shape
        m2 = c.reshape(-1, r1)
        gl = m2.T @ m2
        curr, nxt = gram_eig_and_svd(gl, gr[i + 1], delta)
        rk = curr.shape[1]
        cores[i] = (m2 @ curr).reshape(r0, n, rk)
        c1 = cores[i + 1]
        cores[i + 1] = (nxt @ c1.reshape(c1.shape[0], -1)).reshape(rk, c1.shape[1], c1.shape[2])
    return cores, float(delta)


# ----------------------------------------------------------------------------
# TT-SVD of a dense tensor
# ----------------------------------------------------------------------------
def tt_svd(dense: np.ndarray, eps: float) -> Tuple[Cores, float]:
    """TT-SVD by sequential reshape-and-truncate.

    The reference has no single entry point; this is the composition
    TensorNetwork.svd -> Tensor.svd -> delta_svd + merge(v, s)
    (pytens/algs.py:633-702, :238-274, :735-761; pytens/utils.py:19-100)
    verified in SURVEY.md section 3.3, with the TT-SVD delta of the commented line
    pytens/utils.py:53:  delta = eps / sqrt(d-1) * ||X||_F.
    Returns (cores, delta).
    """
    shape = dense.shape
    d = len(shape)
    delta = eps / math.sqrt(max(d - 1, 1)) * float(np.linalg.norm(dense.ravel()))
    cores = []
    carry = np.asarray(dense, dtype=np.float64).reshape(shape[0], -1)
    r = 1
    for k in range(d - 1):
        mat = carry.reshape(r * shape[k], -1)
        u, s, v, _rem, _ = delta_svd(mat, delta)
        rho = u.shape[1]
        cores.append(u.reshape(r, shape[k], rho))
        carry = s[:, None] * v
        r = rho
    cores.append(carry.reshape(r, shape[d - 1], 1))
    return cores, delta


# ----------------------------------------------------------------------------
# algorithmic work models -- used by bench.py for the
# roofline numerators; kept next to the oracle so both arms use one formula.
# ----------------------------------------------------------------------------
def inner_flops(shape: Sequence[int], ra: Sequence[int], rb: Sequence[int]) -> int:
    """F_inner = sum_k 2 a_{k-1} b_{k-1} n_k b_k + 2 a_{k-1} n_k a_k b_k."""
    a = [1] + list(ra) + [1]
    b = [1] + list(rb) + [1]
    f = 0
    for k, n in enumerate(shape):
        f += 2 * a[k] * b[k] * n * b[k + 1] + 2 * a[k] * n * a[k + 1] * b[k + 1]
    return int(f)


def tt_bytes(shape: Sequence[int], ranks: Sequence[int]) -> int:
    r = [1] + list(ranks) + [1]
    return int(8 * sum(r[k] * n * r[k + 1] for k, n in enumerate(shape)))


def round_flops(shape: Sequence[int], ranks: Sequence[int], new_ranks: Sequence[int]) -> int:
    """FLOP model of the reference rounding algorithm."""
    d = len(shape)
    r = [1] + list(ranks) + [1]
    rho = [1] + list(new_ranks) + [1]
    # the last core's QR caps the bond at min(r, n) (pytens/algs.py:1695-1697)
    r_rq = list(r)
    r_rq[d - 1] = min(r[d - 1], shape[d - 1]) if d > 1 else r[d - 1]
    f = 0.0
    for k in range(d - 1, 0, -1):  # RQ pass on cores d-1 .. 1
        m = shape[k] * r_rq[k + 1]
        c = r[k]
        mm, cc = max(m, c), min(m, c)
        f += 4 * mm * cc * cc - 4 * cc**3 / 3
        f += 2 * r[k - 1] * shape[k - 1] * c * r_rq[k]
    for k in range(0, d - 1):  # forward pass on cores 0 .. d-2
        m = rho[k] * shape[k]
        c = r_rq[k + 1]
        mm, cc = max(m, c), min(m, c)
        f += 4 * mm * cc * cc - 4 * cc**3 / 3 + 22 * cc**3 + 2 * m * cc * rho[k + 1]
        f += 2 * rho[k + 1] * c * shape[k + 1] * r_rq[k + 2]
    return int(f)


def ttsvd_flops(shape: Sequence[int], new_ranks: Sequence[int]) -> int:
    d = len(shape)
    rho = [1] + list(new_ranks) + [1]
    f = 0.0
    for k in range(d - 1):
        m = rho[k] * shape[k]
        c = int(np.prod(shape[k + 1 :]))
        mx, mn = max(m, c), min(m, c)
        f += 2 * mx * mn * mn - 2 * mn**3 / 3 + 22 * mn**3 + 2 * rho[k + 1] * m * c
    return int(f)
