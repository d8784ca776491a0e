"""GPU parity at BASELINE scale: rounding at cfg3's core shape
(n=64, bond 256) and TT-SVD with cfg4's ranks, compared with the numpy oracle
(pytens/algs.py:1841-1903, pytens/utils.py:19-100) -- degenerate (X (+) X) and
generic (decaying spectrum, real truncation) inputs, with and without the
deflation / certificate shortcuts.

Error gate (north_star): |err_ours - err_ref| <= 1e-10 on the relative reconstruction
error, equal truncated ranks.  Where the dense tensor is too large (d=5, 64^5), the
error ||Z - Y|| is taken from the difference TT by QR orthogonalisation on the CPU
(accurate to eps_mach * ||Y||, unlike norms-and-inner-product cancellation).
"""

import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu

ERR_TOL = 1e-10  # north_star
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tt(cores):
    from tensor_networks_b200 import TensorTrain

    return TensorTrain.from_cores(cores)


def _diff_norm(x, y):
    """||X - Y||_F of two TTs from the R factor of the difference train (CPU, QR sweeps)."""
    neg = [c.copy() for c in y]
    neg[0] = -neg[0]
    diff = orc.tt_add([c.copy() for c in x], neg)
    for k in range(len(diff) - 1, 0, -1):
        orc.right_orth(diff, k)
    return float(np.linalg.norm(diff[0]))


def _decaying(shape, parts, r_part, decade, seed):
    """Sum of `parts` random TTs of bond r_part with weights 1, 10^-decade, ... : every
    unfolding has a decaying spectrum with no exact rank deficiency."""
    rng = np.random.default_rng(seed)
    y = None
    for j in range(parts):
        t = orc.rand_tt(shape, [r_part] * (len(shape) - 1), rng)
        t[0] = t[0] * 10.0 ** (-decade * j)
        y = t if y is None else orc.tt_add(y, t)
    return y


def _left_orth_defect(tt):
    worst = 0.0
    for c in tt.cores[:-1]:
        m = c.reshape(-1, c.shape[2])
        g = (m.T @ m).cpu().numpy()
        worst = max(worst, np.abs(g - np.eye(g.shape[0])).max())
    return worst


def _check_round(y, eps, dense_check):
    ref, _ = orc.svd_round([c.copy() for c in y], eps)
    tt = _tt(y).round(eps)
    assert tt.last_round["not_converged"] == 0
    assert tt.ranks() == orc.ranks_of(ref), (tt.ranks(), orc.ranks_of(ref))
    assert _left_orth_defect(tt) < 1e-12
    ny = orc.norm(y)
    if dense_check:
        dense = orc.to_dense(y)
        err = np.linalg.norm(tt.dense() - dense) / np.linalg.norm(dense)
        err_ref = np.linalg.norm(orc.to_dense(ref) - dense) / np.linalg.norm(dense)
    else:
        err = _diff_norm(tt.to_cores(), y) / ny
        err_ref = _diff_norm(ref, y) / ny
    assert abs(err - err_ref) <= ERR_TOL, (err, err_ref)
    assert err <= eps * (1 + 1e-9) + 1e-13
    return tt


def test_round_cfg3_slice_doubled_d4_dense():
    """cfg3 core shape (n=64, bond 256 = X (+) X with X of bond 128), d=4, dense error."""
    rng = np.random.default_rng(2001)
    x = orc.rand_tt([64] * 4, [128] * 3, rng)
    y = orc.tt_add(x, x)
    tt = _check_round(y, 1e-8, dense_check=True)
    assert tt.ranks() == [64, 128, 64]


def test_round_cfg3_slice_doubled_d5():
    """d=5 slice of cfg3: two interior bonds of 256 inside the sweep (p=256 Jacobi / certificate,
    4-panel Cholesky-QR2 at c=256)."""
    rng = np.random.default_rng(2002)
    x = orc.rand_tt([64] * 5, [128] * 4, rng)
    y = orc.tt_add(x, x)
    tt = _check_round(y, 1e-8, dense_check=False)
    assert tt.ranks() == [64, 128, 128, 64]


@pytest.mark.parametrize("eps", [1e-8, 1e-5])
def test_round_cfg3_slice_generic_d4_dense(eps):
    """Bond 256 with a decaying spectrum (8 parts of bond 32, one decade apart): eps really truncates."""
    y = _decaying([64] * 4, parts=8, r_part=32, decade=1.5, seed=77)
    assert orc.ranks_of(y) == [256] * 3
    tt = _check_round(y, eps, dense_check=True)
    assert max(tt.ranks()) < 256  # truncated
    assert tt.last_round["svds"] >= 1


def test_round_cfg3_slice_generic_d5():
    y = _decaying([64] * 5, parts=8, r_part=32, decade=1.5, seed=78)
    _check_round(y, 1e-8, dense_check=False)


_NO_SHORTCUT_SCRIPT = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np
import test_scale_gpu as t
from oracle import tt_oracle as orc
rng = np.random.default_rng(2001)
x = orc.rand_tt([64] * 4, [128] * 3, rng)
tt = t._check_round(orc.tt_add(x, x), 1e-8, dense_check=True)
assert tt.last_round["svds_certified"] == 0 and tt.last_round["bonds_deflated"] == 0, tt.last_round
y = t._decaying([64] * 4, parts=8, r_part=32, decade=1.5, seed=77)
t._check_round(y, 1e-8, dense_check=True)
print("OK", tt.ranks(), tt.last_round)
"""


def test_round_cfg3_slice_without_shortcuts():
    """Same inputs with TTB_DEFLATE=0 TTB_SVD_CERT=0 (the knobs are read once per process):
    every bond goes through the full QR + Jacobi SVD and must give the same ranks / error."""
    env = dict(os.environ, TTB_DEFLATE="0", TTB_SVD_CERT="0")
    code = _NO_SHORTCUT_SCRIPT.format(root=ROOT, tests=os.path.join(ROOT, "tests"))
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "OK" in res.stdout


@pytest.mark.parametrize("noise", [0.0, 1e-13])
def test_ttsvd_cfg4_ranks_16p6_vs_oracle(noise):
    """TT-SVD of a dense 16^6 tensor generated with cfg4's ranks [16, 64, 64, 64, 16], eps=1e-10,
    against orc.tt_svd (pytens/utils.py:19-100 composed as in SURVEY 8c)."""
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(3001)
    gen = orc.rand_tt([16] * 6, [16, 64, 64, 64, 16], rng)
    dense = orc.to_dense(gen)
    if noise:
        dense = dense + noise * np.linalg.norm(dense) / np.sqrt(dense.size) * rng.standard_normal(dense.shape)
    ref, _ = orc.tt_svd(dense, 1e-10)
    tt = TensorTrain.from_dense(dense, 1e-10)
    assert tt.ranks() == orc.ranks_of(ref), (tt.ranks(), orc.ranks_of(ref))
    assert tt.ranks() == [16, 64, 64, 64, 16]
    nd = np.linalg.norm(dense)
    err = np.linalg.norm(tt.dense() - dense) / nd
    err_ref = np.linalg.norm(orc.to_dense(ref) - dense) / nd
    assert abs(err - err_ref) <= ERR_TOL, (err, err_ref)
    assert err <= 1e-10
    assert _left_orth_defect(tt) < 1e-12


def test_ttsvd_16p6_generic_truncation():
    """Dense 16^6 with a decaying TT spectrum: TT-SVD really truncates; ranks equal to the oracle's."""
    from tensor_networks_b200 import TensorTrain

    y = _decaying([16] * 6, parts=6, r_part=8, decade=2.0, seed=31)
    dense = orc.to_dense(y)
    for eps in (1e-4, 1e-8):
        ref, _ = orc.tt_svd(dense, eps)
        tt = TensorTrain.from_dense(dense, eps)
        assert tt.ranks() == orc.ranks_of(ref), (eps, tt.ranks(), orc.ranks_of(ref))
        nd = np.linalg.norm(dense)
        err = np.linalg.norm(tt.dense() - dense) / nd
        err_ref = np.linalg.norm(orc.to_dense(ref) - dense) / nd
        assert abs(err - err_ref) <= ERR_TOL
        assert err <= eps
