"""Gram-SVD rounding on the device against the reference's fixtures and the oracle
(tt_gramsvd_round, pytens/algs.py:1771-1838; reference test: tests/main_test.py:245-262)."""

import copy
import glob
import os

import numpy as np
import pytest

from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _cores(z, prefix):
    out, k = [], 0
    while f"{prefix}{k}" in z:
        out.append(z[f"{prefix}{k}"])
        k += 1
    return orc.as_cores3(out)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "gramsvd_*.npz"))))
def test_gramsvd_matches_reference_fixture(path):
    from tensor_networks_b200 import TensorTrain

    z = np.load(path)
    cores = _cores(z, "in")
    dense = orc.to_dense(cores)
    tt = TensorTrain.from_cores(copy.deepcopy(cores)).gramsvd_round(float(z["eps"]))
    assert tt.ranks() == list(z["ranks_out"])
    out = tt.dense()
    err = np.linalg.norm(out - dense) / np.linalg.norm(dense)
    assert abs(err - float(z["rel_err"])) <= 1e-8  # Gram SVD: sqrt(machine eps) accuracy class
    ref = orc.to_dense(_cores(z, "out"))
    assert np.allclose(out, ref, rtol=0, atol=1e-8 * np.linalg.norm(dense))


def test_gramsvd_like_reference_test():
    """tests/main_test.py:245-262: (X + X) rounded with eps 1e-5 keeps the ranks of X and the values."""
    from tensor_networks_b200 import TensorTrain

    x = TensorTrain.rand([5, 10, 20], [2, 2], seed=4)
    y = x + x
    dense = y.dense()
    z = y.gramsvd_round(1e-5)
    assert z.ranks() == [2, 2]
    assert np.allclose(z.dense(), dense, atol=1e-7 * np.abs(dense).max(), rtol=1e-7)


@pytest.mark.parametrize("seed,eps", [(1, 1e-3), (2, 1e-5), (3, 1e-6)])
def test_gramsvd_matches_oracle_larger(seed, eps):
    """Bonds up to 96 with a decaying spectrum: ranks equal the oracle's, error within the class tolerance."""
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(700 + seed)
    shape = [10] * 5
    y = orc.rand_tt(shape, [24] * 4, rng)
    for j in range(1, 4):
        zt = orc.rand_tt(shape, [24] * 4, rng)
        zt[0] = zt[0] * 10.0 ** (-2 * j)
        y = orc.tt_add(y, zt)
    dense = orc.to_dense(y)
    ref, delta_ref = orc.gramsvd_round(copy.deepcopy(y), eps)
    tt = TensorTrain.from_cores(copy.deepcopy(y)).gramsvd_round(eps)
    assert tt.ranks() == orc.ranks_of(ref)
    assert abs(tt.last_gramsvd["delta"] - delta_ref) <= 1e-12 * delta_ref
    err = np.linalg.norm(tt.dense() - dense) / np.linalg.norm(dense)
    err_ref = np.linalg.norm(orc.to_dense(ref) - dense) / np.linalg.norm(dense)
    assert abs(err - err_ref) <= 1e-8
    assert err <= 2 * eps


def test_gramsvd_network_api():
    """algs.tt_gramsvd_round on a TensorNetwork mutates it in place and returns it, like the reference."""
    from tensor_networks_b200 import algs

    np.random.seed(9)
    idx = [algs.Index(f"x{i}", n) for i, n in enumerate([6, 7, 8, 5])]
    a = algs.TensorNetwork.rand_tt(idx, [3, 4, 3])
    s = a + a
    dense = s.contract().value
    out = algs.tt_gramsvd_round(s, 1e-6)
    assert out is s
    assert s.ranks() == [3, 4, 3]
    assert np.allclose(s.contract().value, dense, atol=1e-7 * np.abs(dense).max())


def test_gram_eig_and_svd_host_api():
    from tensor_networks_b200 import algs

    rng = np.random.default_rng(12)
    a = rng.standard_normal((40, 12)) @ np.diag(10.0 ** -np.arange(12)) 
    b = rng.standard_normal((12, 50))
    gl, gr = a.T @ a, b @ b.T
    delta = 1e-4 * np.linalg.norm(a @ b)
    c_ref, n_ref = orc.gram_eig_and_svd(gl, gr, delta)
    c, n = algs.gram_eig_and_svd(gl, gr, delta)
    assert c.shape == c_ref.shape and n.shape == n_ref.shape
    # the factors are unique up to an orthogonal mixing; the rounded product a c n b is not
    assert np.allclose(a @ c @ n @ b, a @ c_ref @ n_ref @ b, atol=1e-8 * np.linalg.norm(a @ b))


@pytest.mark.parametrize("count,p,decades", [(1, 1, 0.0), (3, 2, 1.0), (5, 37, 0.5), (19, 128, 0.13), (2, 256, 0.07)])
def test_gram_eig_batched_matches_eigh(count, p, decades):
    """ttb_gram_eig_batched_f64 (one cluster-Jacobi launch for the whole stack) against numpy's eigh plus the
    reference's rounding rule (pytens/algs.py:1727-1749): eigenvalues, A A^T = V diag(e12^2) V^T and B from A."""
    import torch

    from tensor_networks_b200.gramsvd import gram_eig_batched_dev

    rng = np.random.default_rng(100 + p)
    gs = []
    for _ in range(count):
        x = rng.standard_normal((p + 5, p)) * 10.0 ** (-decades * np.arange(p))[None, :]  # graded columns: Gram spans 2x the decades
        q = np.linalg.qr(rng.standard_normal((p, p)))[0]
        gs.append(q @ (x.T @ x) @ q.T)
    g = np.stack(gs)
    a, b, eig = (t.cpu().numpy() for t in gram_eig_batched_dev(torch.from_numpy(g).cuda()))
    for i in range(count):
        w, v = np.linalg.eigh(g[i])
        w = np.abs(w)[::-1]
        assert np.all(np.abs(eig[i] - w) <= 1e-13 * w[0]), np.max(np.abs(eig[i] - w)) / w[0]
        e12 = orc.round_sqrt_eigs(w)
        # the device values pass through the same rounding grid: identical up to a grid step at rounding boundaries
        step = 10.0 ** np.ceil(np.log10(e12.max() * 1e-8 + 1e-15))
        got12 = np.sqrt(np.sum(a[i] ** 2, axis=0))  # column norms of V diag(e12)
        # (the square root of an eigenvalue at the roundoff level of eigh, ~1e-14 of the largest, is not determined:
        # two backward-stable solvers differ there by several grid steps -- compared above 1e-10 of the largest only)
        big = w > 1e-10 * w[0]
        assert np.all(np.abs(got12 - e12)[big] <= 1.01 * step), np.max(np.abs(got12 - e12)[big]) / step
        assert np.all(got12[~big] <= 1e-5 * e12[0])
        # A A^T reproduces G to the rounding of the square roots
        assert np.linalg.norm(a[i] @ a[i].T - g[i]) <= 4e-8 * np.linalg.norm(g[i]) * np.sqrt(p)
        nz = got12 > 0
        assert np.allclose(b[i][:, nz], a[i][:, nz] / got12[nz] ** 2, rtol=1e-12, atol=0.0)
        assert np.all(b[i][:, ~nz] == 0.0)
        # columns are orthogonal (eigenvectors) wherever they are not rounded away
        vn = a[i][:, nz] / got12[nz]
        assert np.abs(vn.T @ vn - np.eye(vn.shape[1])).max() <= 1e-10
