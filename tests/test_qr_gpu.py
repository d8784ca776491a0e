"""GPU tests of the tall-skinny orthogonalisation (TSQR Householder panels + BCGS with
DGKS-controlled re-orthogonalisation) against numpy's Householder QR, including the
rank-deficient inputs that TT sums (X + X) produce."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(a_np, orth_tol=1e-13):
    from tensor_networks_b200.utils import orth_rows_dev

    a = torch.from_numpy(a_np).cuda()
    c, m = a.shape
    q, r = orth_rows_dev(a.clone())
    q, r = q.cpu().numpy(), r.cpu().numpy()
    k = min(c, m)
    g = q[:k] @ q[:k].T
    assert np.abs(g - np.eye(k)).max() < orth_tol
    if c > m:
        assert np.abs(q[k:]).max() == 0.0  # zero-padding branch (pytens/algs.py:1679-1685)
    assert np.linalg.norm(r.T @ q - a_np) <= 1e-14 * max(1.0, np.linalg.norm(a_np)) * np.sqrt(c)
    assert np.abs(np.tril(r[:k, :k], -1)).max() == 0.0
    # same R up to row signs as LAPACK's QR of the transpose (np.linalg.qr(val.T), algs.py:1678)
    if np.linalg.matrix_rank(a_np) == k and c <= m:
        r_ref = np.linalg.qr(a_np.T, mode="r")
        assert np.allclose(np.abs(np.diag(r)), np.abs(np.diag(r_ref)), rtol=1e-9)


@pytest.mark.parametrize("c,m", [(1, 1), (1, 50), (7, 5), (32, 256), (33, 257), (100, 64), (256, 4096), (96, 20000), (300, 300)])
def test_orth_rows_random(c, m):
    rng = np.random.default_rng(c * 1000 + m)
    _check(rng.standard_normal((c, m)))


@pytest.mark.parametrize("c,m,rank", [(256, 4096, 64), (1024, 4096, 64), (512, 2048, 3), (64, 1000, 1)])
def test_orth_rows_rank_deficient(c, m, rank):
    rng = np.random.default_rng(1)
    a = rng.standard_normal((c, rank)) @ rng.standard_normal((rank, m)) / np.sqrt(rank * m)
    _check(a)


def test_orth_rows_adversarial():
    rng = np.random.default_rng(2)
    a = rng.standard_normal((128, 2000))
    a[40] = a[35]  # duplicate inside a panel
    a[70] = a[3]  # duplicate across panels
    a[100:110] = 0.0  # zero vectors
    a[120] = 1e-200 * a[5]  # tiny copy
    _check(a)
    b = rng.standard_normal((64, 512))
    _check(np.vstack([b, b]))  # X (+) X like: second half duplicates the first
    g = np.diag(np.logspace(0, -18, 96)) @ rng.standard_normal((96, 700))  # graded rows
    _check(g)


@pytest.mark.parametrize("c,m", [(100, 3001), (37, 2049), (130, 18944), (64, 18945), (64, 2048), (200, 8190), (65, 16384)])
def test_orth_rows_fused_panel_shapes(c, m):
    """Shapes on both sides of the fused cooperative panel kernel's range (2048 <= m <= 148 * 128), with
    odd row lengths (scalar slab loads, partial last slab) and panel widths that are not multiples of 64."""
    rng = np.random.default_rng(c * 7 + m)
    _check(rng.standard_normal((c, m)))


def test_orth_rows_fused_panel_decisions():
    """Inside the fused kernel's range: a duplicated block (deflation is NOT active through this entry
    point, so the dependent panel breaks down and goes to the Householder path), an ill-conditioned
    panel (declined on the device, panel left untouched) and a badly scaled one."""
    rng = np.random.default_rng(5)
    b = rng.standard_normal((64, 4096))
    _check(np.vstack([b, b, rng.standard_normal((20, 4096))]))
    mix = rng.standard_normal((64, 64))
    mix[:, -1] = mix[:, 0] * (1 + 1e-7)  # two nearly parallel directions: cond ~ 1e7
    _check(np.vstack([mix @ rng.standard_normal((64, 4096)), rng.standard_normal((64, 4096))]), orth_tol=1e-12)
    s = rng.standard_normal((128, 4100)) * np.logspace(0, -12, 128)[:, None]
    _check(s)


def test_orth_rows_fused_repeatable():
    """Plan replay and graph capture on the fused path: the same call five times gives bit-identical Q and R."""
    from tensor_networks_b200.utils import orth_rows_dev

    rng = np.random.default_rng(8)
    a = torch.from_numpy(rng.standard_normal((192, 8192))).cuda()
    outs = []
    for _ in range(5):
        q, r = orth_rows_dev(a.clone())
        outs.append((q.clone(), r.clone()))
    for q, r in outs[1:]:
        assert torch.equal(q, outs[0][0]) and torch.equal(r, outs[0][1])
