"""GPU parity for TT rounding (north_star gate: reconstruction error within 1e-10
relative of the reference's, equal truncated ranks for the same eps)."""

import copy

import numpy as np
import pytest
import torch

from conftest import golden_files, load_cores
from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu

ERR_TOL = 1e-10  # north_star


def _tt(cores):
    from tensor_networks_b200 import TensorTrain

    return TensorTrain.from_cores(cores)


def _left_orth_defect(tt):
    worst = 0.0
    for c in tt.cores[:-1]:
        m = c.reshape(-1, c.shape[2])
        g = (m.T @ m).cpu().numpy()
        worst = max(worst, np.abs(g - np.eye(g.shape[0])).max())
    return worst


@pytest.mark.parametrize("path", golden_files("round"))
def test_round_golden(path):
    z = np.load(path)
    cores = load_cores(z, "in")
    dense = orc.to_dense(orc.as_cores3(cores))
    tt = _tt(cores)
    out = tt.round(float(z["eps"]))
    assert out is tt  # in place, returns the same object (pytens/algs.py:1903)
    assert tt.ranks() == list(z["ranks_out"]), (tt.ranks(), list(z["ranks_out"]))
    got = tt.dense()
    err = np.linalg.norm(got - dense) / np.linalg.norm(dense)
    assert abs(err - float(z["rel_err"])) <= ERR_TOL, (err, float(z["rel_err"]))
    assert err <= float(z["eps"]) * (1 + 1e-9) + 1e-13
    ref_dense = orc.to_dense(orc.as_cores3(load_cores(z, "out")))
    assert np.linalg.norm(got - ref_dense) <= max(10 * float(z["rel_err"]), 1e-11) * np.linalg.norm(dense) * 2
    assert _left_orth_defect(tt) < 1e-12
    assert tt.last_round["not_converged"] == 0


@pytest.mark.parametrize("path", golden_files("right_orth"))
def test_right_orth_golden(path):
    z = np.load(path)
    cores = load_cores(z, "in")
    d = len(cores)
    dense = orc.to_dense(orc.as_cores3(cores))
    tt = _tt(cores)
    tt.right_orth(d - 1)
    ref = orc.as_cores3(load_cores(z, "after_last_"))
    assert [tuple(c.shape) for c in tt.cores] == [c.shape for c in ref]
    for j in range(d - 2, 0, -1):
        tt.right_orth(j)
    ref = orc.as_cores3(load_cores(z, "after_all_"))
    assert [tuple(c.shape) for c in tt.cores] == [c.shape for c in ref]
    # reference test_right_orthogonalization (tests/main_test.py:200-224)
    for k in range(1, d):
        m = tt.cores[k].reshape(tt.cores[k].shape[0], -1).cpu().numpy()
        g = m @ m.T
        nz = np.abs(np.diag(g)) > 0.5  # zero-padded rows (algs.py:1679-1685) stay zero
        assert int(nz.sum()) == min(m.shape)
        assert np.allclose(g[np.ix_(nz, nz)], np.eye(int(nz.sum())), atol=1e-13)
        assert np.abs(g[~nz]).max(initial=0.0) < 1e-25
    assert np.allclose(tt.dense(), dense, rtol=1e-12, atol=1e-12 * np.abs(dense).max())


@pytest.mark.parametrize("path", golden_files("delta_svd"))
def test_delta_svd_golden(path):
    from tensor_networks_b200.utils import delta_svd

    z = np.load(path)
    t = delta_svd(z["mat"], float(z["delta_in"]), bool(z["with_normalizing"]))
    assert len(t.s) == len(z["s"])
    smax = z["s"][0]
    assert np.allclose(t.s, z["s"], rtol=1e-9, atol=1e-13 * smax)
    assert abs(t.remaining_delta - float(z["remaining_delta"])) <= 1e-9 * max(1e-300, abs(float(z["remaining_delta"]))) + 1e-13 * smax
    if bool(z["with_normalizing"]):
        assert abs(t.delta - float(z["delta_out"])) <= 1e-12 * t.delta
    else:
        assert t.delta is None
    ref = (z["u"] * z["s"]) @ z["v"]
    got = (t.u * t.s) @ t.v
    assert np.allclose(got, ref, atol=1e-9 * smax)
    assert np.allclose(t.u.T @ t.u, np.eye(len(t.s)), atol=1e-12)


@pytest.mark.parametrize(
    "shape,ranks,eps",
    [
        ([16] * 6, [12] * 5, 1e-8),
        ([9, 4, 11, 6, 5, 7], [5, 9, 13, 7, 3], 1e-10),
        ([40, 3, 40], [37, 33], 1e-8),  # wide and tall unfoldings, odd sizes
        ([2] * 12, [2, 4, 8, 16, 32, 40, 32, 16, 8, 4, 2], 1e-6),
        ([3, 70], [2], 1e-8),
    ],
)
def test_round_vs_oracle_doubled(shape, ranks, eps):
    rng = np.random.default_rng(99)
    x = orc.rand_tt(shape, ranks, rng)
    y = orc.tt_add(x, x)
    ref, _ = orc.svd_round(copy.deepcopy(y), eps)
    tt = _tt(y).round(eps)
    assert tt.ranks() == orc.ranks_of(ref)
    if np.prod(shape) <= 4_000_000:
        dense = orc.to_dense(y)
        err = np.linalg.norm(tt.dense() - dense) / np.linalg.norm(dense)
        err_ref = np.linalg.norm(orc.to_dense(ref) - dense) / np.linalg.norm(dense)
        assert abs(err - err_ref) <= ERR_TOL
    assert _left_orth_defect(tt) < 1e-12


@pytest.mark.parametrize("eps", [1e-1, 1e-2, 1e-4])
def test_round_genuine_truncation(eps):
    """Decaying spectrum: sum of TTs with weights 1, 1e-1, 1e-2 ... (ranks really truncate)."""
    rng = np.random.default_rng(5)
    shape = [6] * 7
    y = orc.rand_tt(shape, [3] * 6, rng)
    for j in range(1, 5):
        zt = orc.rand_tt(shape, [2] * 6, rng)
        zt[0] = zt[0] * 10.0 ** (-j)
        y = orc.tt_add(y, zt)
    dense = orc.to_dense(y)
    ref, _ = orc.svd_round(copy.deepcopy(y), eps)
    tt = _tt(y).round(eps)
    err = np.linalg.norm(tt.dense() - dense) / np.linalg.norm(dense)
    err_ref = np.linalg.norm(orc.to_dense(ref) - dense) / np.linalg.norm(dense)
    assert tt.ranks() == orc.ranks_of(ref), (tt.ranks(), orc.ranks_of(ref))
    assert abs(err - err_ref) <= ERR_TOL
    assert err <= eps


def test_round_max_rank_and_rank_one():
    rng = np.random.default_rng(8)
    x = orc.rand_tt([5] * 6, [6] * 5, rng)
    dense = orc.to_dense(x)
    tt = _tt(x).round(1e-12, max_rank=3)
    assert tt.ranks() == [3] * 5
    # huge eps: everything truncated -> rank clamps to 1 (pytens/utils.py:84)
    tt1 = _tt(x).round(10.0)
    assert tt1.ranks() == [1] * 5
    assert np.isfinite(tt1.dense()).all()
    # a rank-1 TT is unchanged by rounding
    r1 = orc.rand_tt([4] * 5, [1] * 4, rng)
    out = _tt(r1).round(1e-10)
    assert out.ranks() == [1] * 4
    assert np.allclose(out.dense(), orc.to_dense(r1), rtol=1e-12, atol=1e-14)
    assert dense.shape == (5,) * 6


def test_round_medium_rank_vs_oracle():
    """d=8, n=32, 64 -> doubled 128: multi-panel QR, multi-block Jacobi, split-K GEMMs."""
    from tensor_networks_b200 import TensorTrain

    d, n, r = 8, 32, 64
    x = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2001)
    y = x + x
    ycores = y.to_cores()
    ref, _ = orc.svd_round([c.copy() for c in ycores], 1e-8)
    ny = y.norm()
    z = y.clone().round(1e-8)
    assert z.ranks() == orc.ranks_of(ref), (z.ranks(), orc.ranks_of(ref))
    assert z.ranks() == [min(n, r)] + [r] * (d - 3) + [min(n, r)]
    assert _left_orth_defect(z) < 1e-12
    nz = z.norm()
    assert abs(nz - ny) <= 1e-10 * ny
    cosang = float(z.inner(y)) / (ny * nz)
    assert abs(cosang - 1.0) < 1e-12
    # idempotence: rounding the rounded TT keeps ranks and the tensor
    z2 = z.clone().round(1e-8)
    assert z2.ranks() == z.ranks()
    assert abs(float(z2.inner(z)) / (z2.norm() * nz) - 1.0) < 1e-12


@pytest.mark.parametrize("scale", [1e-6, 1e-9, 1e-11])
def test_round_under_cancellation_matches_reference(scale):
    """A formal sum with heavy cancellation, x + (-x) + scale * z (the shape of a GMRES residual b - A x): the rows that
    deflation drops are exact dependencies, so ranks and accuracy stay those of the reference's LAPACK path -- also
    where the reference itself degrades (scale = 1e-11: both keep noise ranks and lose five digits)."""
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(7)
    shape = [8] * 6
    x = orc.rand_tt(shape, [6] * 5, rng)
    z = orc.rand_tt(shape, [4] * 5, rng)
    xm = [c.copy() for c in x]
    xm[0] = -xm[0]
    zs = [c.copy() for c in z]
    zs[0] = zs[0] * scale
    y = orc.tt_add(orc.tt_add(x, xm), zs)
    want = orc.to_dense(zs)
    ref_cores, _ = orc.svd_round([c.copy() for c in y], 1e-6)
    ref_err = np.linalg.norm(orc.to_dense(ref_cores) - want) / np.linalg.norm(want)
    t = TensorTrain.from_cores(y)
    t.round(1e-6)
    err = np.linalg.norm(t.dense() - want) / np.linalg.norm(want)
    assert t.ranks() == orc.ranks_of(ref_cores)
    assert err <= 2.0 * ref_err + 1e-10
