"""On-device SVD (cluster block-Jacobi kernel) and truncation paths against numpy / the oracle.

Shapes are chosen so that every launch shape of the single-launch cluster kernel is hit
(1, 2, 3, 4, 5, 8 CTAs per cluster, padded last blocks), plus the multi-launch fallback
(p > 256) and the tall / wide-LQ / wide-direct paths of trunc_svd.
"""

import copy

import numpy as np
import pytest

from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _graded(m, n, decay, rng):
    """Random m x n matrix with singular values 1, decay, decay^2, ... (graded spectrum)."""
    p = min(m, n)
    u, _ = np.linalg.qr(rng.standard_normal((m, p)))
    v, _ = np.linalg.qr(rng.standard_normal((n, p)))
    s = decay ** np.arange(p)
    return (u * s) @ v.T, s


@pytest.mark.parametrize(
    "m,n,decay",
    [
        (24, 24, 0.5),      # one CTA
        (40, 40, 0.7),      # cluster of 2, padded block
        (96, 96, 0.8),      # cluster of 3
        (100, 70, 0.75),    # tall, cluster of 3 (p = 70)
        (130, 200, 0.85),   # wide direct (c < 2m), cluster of 5
        (256, 256, 0.9),    # cluster of 8
        (300, 64, 0.6),     # tall: QR first
        (64, 300, 0.6),     # wide LQ
        (1000, 250, 0.9),   # tall, 4 panels
        (300, 300, 0.93),   # p > 256: multi-launch fallback
    ],
)
def test_delta_svd_graded_vs_numpy(m, n, decay):
    from tensor_networks_b200.utils import delta_svd_dev

    rng = np.random.default_rng(m * 1000 + n)
    a, s_true = _graded(m, n, decay, rng)
    u, s, svt, info = delta_svd_dev(torch.from_numpy(a).cuda(), 0.0)
    u, s, svt = u.cpu().numpy(), s.cpu().numpy(), svt.cpu().numpy()
    s_np = np.linalg.svd(a, compute_uv=False)
    assert info["rank"] == min(m, n)
    big = s_np > 1e-9 * s_np[0]
    # LAPACK (and the construction of `a` itself) is only accurate to eps * sigma_max in the absolute
    # sense, so that is the yardstick; the large values also agree relatively
    assert np.max(np.abs(s - s_np)) < 2e-13 * s_np[0]  # rotations accumulate a few hundred ulp
    top = s_np > 1e-4 * s_np[0]
    assert np.max(np.abs(s[top] - s_np[top]) / s_np[top]) < 1e-10
    assert np.linalg.norm(u @ svt - a) <= 1e-13 * np.linalg.norm(a)
    k = int(big.sum())
    g = u[:, :k].T @ u[:, :k]
    assert np.max(np.abs(g - np.eye(k))) < 1e-12


@pytest.mark.parametrize("m,c,cond", [(16, 131072, 1e2), (12, 65536 + 16 * 7, 1e4), (16, 262144, 1e9), (5, 98304, 1.0)])
def test_delta_svd_skinny_lq(m, c, cond):
    """Very wide matrices with at most 16 rows take the three-pass skinny LQ (Gram, fused L1^-1 apply + second Gram,
    carry with L2^-1 folded in); cond = 1e9 makes the host Cholesky decline, so the generic panel path is checked
    on the same shape.  Singular values, reconstruction and orthonormality of U against numpy."""
    from tensor_networks_b200.utils import delta_svd_dev

    rng = np.random.default_rng(m + c)
    u, _ = np.linalg.qr(rng.standard_normal((m, m)))
    s_true = np.logspace(0.0, -np.log10(cond), m) if cond > 1.0 else np.ones(m)
    a = (u * s_true) @ (rng.standard_normal((m, c)) / np.sqrt(c))
    U, s, svt, info = delta_svd_dev(torch.from_numpy(a).cuda(), 0.0)
    U, s, svt = U.cpu().numpy(), s.cpu().numpy(), svt.cpu().numpy()
    s_np = np.linalg.svd(a, compute_uv=False)
    assert info["rank"] == m
    assert np.max(np.abs(s - s_np)) < 1e-12 * s_np[0]
    assert np.linalg.norm(U @ svt - a) <= 1e-12 * np.linalg.norm(a)
    assert np.max(np.abs(U.T @ U - np.eye(m))) < 1e-11
    # truncation through the same path: drop the smaller half of the spectrum
    if cond > 1.0:
        delta = 0.5 * (s_np[m // 2 - 1] + s_np[m // 2]) if m >= 4 else 0.0
        tail = np.sqrt(np.sum(s_np[m // 2:] ** 2))
        U2, s2, svt2, info2 = delta_svd_dev(torch.from_numpy(a).cuda(), float(tail * (1 + 1e-9)))
        assert info2["rank"] == m // 2
        err = np.linalg.norm(U2.cpu().numpy() @ svt2.cpu().numpy() - a)
        assert abs(err - tail) <= 1e-9 * np.linalg.norm(a)


def test_delta_svd_width_limit():
    """The Jacobi kernels hold eight rows of [X | J] in shared memory: the documented limit is 1664 singular
    values for tall / very wide inputs.  The boundary case works (multi-launch path, rows of 3332 columns),
    one block more is refused with an explanation instead of a failed launch."""
    from tensor_networks_b200.utils import delta_svd_dev

    rng = np.random.default_rng(3)
    m, p = 1800, 1664
    a = rng.standard_normal((m, 40)) @ rng.standard_normal((40, p)) + 1e-7 * rng.standard_normal((m, p))
    u, s, svt, info = delta_svd_dev(torch.from_numpy(a).cuda(), 1e-4 * np.linalg.norm(a))
    assert info["rank"] == 40
    s_np = np.linalg.svd(a, compute_uv=False)[:40]
    assert np.max(np.abs(s.cpu().numpy() - s_np) / s_np) < 1e-10
    rec = (u @ svt).cpu().numpy()
    assert np.linalg.norm(rec - a) <= 2e-4 * np.linalg.norm(a)
    b = rng.standard_normal((1700, 1672))
    with pytest.raises(Exception, match="wider than the on-chip Jacobi"):
        delta_svd_dev(torch.from_numpy(b).cuda(), 0.0)


@pytest.mark.parametrize("m,n,rank", [(200, 96, 40), (96, 200, 40), (512, 128, 17), (64, 64, 1)])
def test_delta_svd_rank_deficient(m, n, rank):
    """Exactly rank-deficient input: the dropped part is roundoff, the kept part is exact."""
    from tensor_networks_b200.utils import delta_svd_dev

    rng = np.random.default_rng(7 + m + n)
    a = rng.standard_normal((m, rank)) @ rng.standard_normal((rank, n))
    u, s, svt, info = delta_svd_dev(torch.from_numpy(a).cuda(), 1e-10 * np.linalg.norm(a))
    assert info["rank"] == rank
    s_np = np.linalg.svd(a, compute_uv=False)[:rank]
    assert np.max(np.abs(s.cpu().numpy() - s_np) / s_np) < 1e-11
    rec = (u @ svt).cpu().numpy()
    assert np.linalg.norm(rec - a) <= 1e-12 * np.linalg.norm(a)


def test_round_large_rank_genuine_truncation():
    """Ranks 72 with a decaying spectrum: the SVD really truncates (no deflation, no certificate)."""
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(11)
    shape = [12] * 5
    y = orc.rand_tt(shape, [24] * 4, rng)
    for j in range(1, 5):
        zt = orc.rand_tt(shape, [12] * 4, rng)
        zt[0] = zt[0] * 10.0 ** (-2 * j)
        y = orc.tt_add(y, zt)
    dense = orc.to_dense(y)
    for eps in (1e-3, 1e-5, 1e-7):
        ref, _ = orc.svd_round(copy.deepcopy(y), eps)
        tt = TensorTrain.from_cores(copy.deepcopy(y)).round(eps)
        assert tt.ranks() == orc.ranks_of(ref), (eps, tt.ranks(), orc.ranks_of(ref))
        err = np.linalg.norm(tt.dense() - dense) / np.linalg.norm(dense)
        err_ref = np.linalg.norm(orc.to_dense(ref) - dense) / np.linalg.norm(dense)
        assert abs(err - err_ref) <= 1e-10
        assert err <= eps


def test_round_full_rank_is_identity():
    """Well-conditioned full-rank TT and a tiny eps: nothing may be truncated (certificate path)."""
    from tensor_networks_b200 import TensorTrain

    x = TensorTrain.rand([20] * 5, [40] * 4, seed=77)
    ref, _ = orc.svd_round([c.copy() for c in x.to_cores()], 1e-12)
    z = x.clone().round(1e-12)
    assert z.ranks() == orc.ranks_of(ref)
    assert z.last_round["svds_certified"] >= 1 and z.last_round["bonds_deflated"] == 0
    nx, nz = x.norm(), z.norm()
    assert abs(nx - nz) <= 1e-12 * nx
    assert abs(float(z.inner(x)) / (nx * nz) - 1.0) < 1e-12


@pytest.mark.parametrize("r", [17, 33, 64, 65, 100, 128])
def test_round_certificate_sizes(r):
    """The no-truncation certificate (blocked triangular inverse, padded to 128) at bond ranks on
    both sides of its 32- and 64-wide block boundaries."""
    from tensor_networks_b200 import TensorTrain

    x = TensorTrain.rand([130, 5, 5, 130], [r, r, r], seed=300 + r)
    ref, _ = orc.svd_round([c.copy() for c in x.to_cores()], 1e-12)
    z = x.clone().round(1e-12)
    assert z.ranks() == orc.ranks_of(ref)
    assert z.ranks() == [r, r, r]
    assert z.last_round["svds_certified"] >= 1
    nx, nz = x.norm(), z.norm()
    assert abs(nx - nz) <= 1e-12 * nx
    assert abs(float(z.inner(x)) / (nx * nz) - 1.0) < 1e-12


def test_round_certificate_rejects_small_sigma():
    """Full numerical rank but sigma_min below delta: the certificate must NOT fire and the SVD truncates."""
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(5)
    shape = [40, 6, 40]
    big = orc.rand_tt(shape, [20, 20], rng)
    small = orc.rand_tt(shape, [12, 12], rng)
    small[0] = small[0] * 1e-7
    y = orc.tt_add(big, small)
    ref, _ = orc.svd_round(copy.deepcopy(y), 1e-4)
    tt = TensorTrain.from_cores(copy.deepcopy(y)).round(1e-4)
    assert tt.ranks() == orc.ranks_of(ref)
    assert tt.ranks() == [20, 20]
    assert tt.last_round["svds_certified"] == 0


def test_round_partial_deflation():
    """X (+) X (+) noise-free third term of different rank: some panels deflate, some do not."""
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(3)
    shape = [12] * 5
    x = orc.rand_tt(shape, [70] * 4, rng)
    w = orc.rand_tt(shape, [20] * 4, rng)
    y = orc.tt_add(orc.tt_add(x, w), x)  # bonds 160 in panels of 64: [X | X,W,X' | X'] -> only the last panel is dependent
    ref, _ = orc.svd_round(copy.deepcopy(y), 1e-9)
    tt = TensorTrain.from_cores(copy.deepcopy(y)).round(1e-9)
    assert tt.ranks() == orc.ranks_of(ref)
    assert tt.last_round["bonds_deflated"] >= 1
    dense = orc.to_dense(y)
    err = np.linalg.norm(tt.dense() - dense) / np.linalg.norm(dense)
    err_ref = np.linalg.norm(orc.to_dense(ref) - dense) / np.linalg.norm(dense)
    assert abs(err - err_ref) <= 1e-10


def test_round_plan_replay_contradiction():
    """orth_rows replays the decisions recorded for the previous call of the same shape; data that
    contradicts the plan must trigger the synchronous redo and still give the reference result.
    Alternate between inputs WITH exact rank deficiency (panels deflate) and WITHOUT (nothing
    deflates) on identical shapes, several times."""
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(21)
    shape = [16] * 5
    x = orc.rand_tt(shape, [72] * 4, rng)
    doubled = orc.tt_add(x, x)                      # bonds 144, true ranks 72 (first/last 16)
    generic = orc.rand_tt(shape, [144] * 4, rng)    # same shapes, full rank
    for trial in range(3):
        for y in (doubled, generic):
            ref, _ = orc.svd_round(copy.deepcopy(y), 1e-9)
            tt = TensorTrain.from_cores(copy.deepcopy(y)).round(1e-9)
            assert tt.ranks() == orc.ranks_of(ref), (trial, tt.ranks(), orc.ranks_of(ref))
            ny = np.sqrt(orc.inner(y, y))
            z = TensorTrain.from_cores(copy.deepcopy(y))
            nz = tt.norm()
            assert abs(nz - ny) <= 1e-9 * ny
            assert abs(float(tt.inner(z)) / (ny * nz) - 1.0) < 1e-12


@pytest.mark.parametrize("seed", range(10))
def test_round_randomised_structures(seed):
    """Random mixtures X1 (+) X2 (+) X1 (+) ... with bonds of 70-200: several Cholesky-QR panels per
    core with deflated panels in arbitrary positions, compaction moves, the bulk test, plan replay
    across cores, certificate on some cores and real truncation on others."""
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(1000 + seed)
    d = int(rng.integers(3, 6))
    shape = [int(rng.integers(6, 15)) for _ in range(d)]
    while np.prod(shape) > 1_500_000:
        shape[int(np.argmax(shape))] -= 2
    parts = []
    for _ in range(int(rng.integers(2, 4))):
        r = [int(rng.integers(8, 60)) for _ in range(d - 1)]
        parts.append(orc.rand_tt(shape, r, rng))
    order = [0, 1, 0] + ([2, 1] if len(parts) > 2 else [])
    rng.shuffle(order)
    y = None
    for j, idx in enumerate(order):
        t = copy.deepcopy(parts[idx])
        t[0] = t[0] * float(10.0 ** (-rng.integers(0, 4)))
        y = t if y is None else orc.tt_add(y, t)
    eps = float(10.0 ** (-rng.integers(5, 10)))
    ref, _ = orc.svd_round(copy.deepcopy(y), eps)
    tt = TensorTrain.from_cores(copy.deepcopy(y)).round(eps)
    dense = orc.to_dense(y)
    err = np.linalg.norm(tt.dense() - dense) / np.linalg.norm(dense)
    err_ref = np.linalg.norm(orc.to_dense(ref) - dense) / np.linalg.norm(dense)
    assert tt.ranks() == orc.ranks_of(ref), (shape, eps, tt.ranks(), orc.ranks_of(ref))
    assert abs(err - err_ref) <= 1e-10
    assert err <= eps * 1.0000001 + 1e-13


def test_round_zero_tensor_and_unit_modes():
    """Edge cases the reference handles implicitly: an identically zero TT (delta = 0, every rank clamps to 1)
    and modes of size 1."""
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(1)
    x = orc.rand_tt([4, 5, 6], [3, 3], rng)
    x[1] = x[1] * 0.0
    ref, _ = orc.svd_round(copy.deepcopy(x), 1e-8)
    t = TensorTrain.from_cores(copy.deepcopy(x)).round(1e-8)
    assert t.ranks() == orc.ranks_of(ref) == [1, 1]
    assert float(t.norm()) == 0.0
    x = orc.rand_tt([1, 7, 1, 5], [1, 3, 3], rng)
    y = orc.tt_add(x, x)
    ref, _ = orc.svd_round(copy.deepcopy(y), 1e-10)
    t = TensorTrain.from_cores(copy.deepcopy(y)).round(1e-10)
    assert t.ranks() == orc.ranks_of(ref) == [1, 3, 3]
    assert np.abs(t.dense() - orc.to_dense(y)).max() <= 1e-13 * np.abs(orc.to_dense(y)).max()
    t2 = TensorTrain.from_cores(copy.deepcopy(y)).gramsvd_round(1e-6)
    r2, _ = orc.gramsvd_round(copy.deepcopy(y), 1e-6)
    assert t2.ranks() == orc.ranks_of(r2) == [1, 3, 3]
