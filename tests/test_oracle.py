"""CPU tests: pin the numpy oracle to the reference's golden vectors.

The fixtures under tests/golden/ were produced by the unmodified reference
(oracle/make_golden.py).  When /root/reference is present (build container)
the oracle is additionally compared with the live reference side by side.
"""

import copy

import numpy as np
import pytest

from conftest import golden_files, load_cores
from oracle import refshim
from oracle import tt_oracle as orc


@pytest.mark.parametrize("path", golden_files("inner"))
def test_inner_matches_reference(path):
    z = np.load(path)
    a = orc.as_cores3(load_cores(z, "a"))
    b = orc.as_cores3(load_cores(z, "b"))
    val = orc.inner(a, b)
    assert isinstance(val, np.ndarray) and val.shape == () and val.dtype == np.float64
    ref = float(z["inner"])
    # north_star gate: 1e-12 relative
    assert abs(float(val) - ref) <= 1e-12 * abs(ref)
    assert abs(orc.norm(a) - float(z["norm_a"])) <= 1e-12 * float(z["norm_a"])
    assert abs(orc.norm(b) - float(z["norm_b"])) <= 1e-12 * float(z["norm_b"])
    # test_inner of the reference (tests/main_test.py:119-126): against dense
    if np.prod(z["shape"]) <= 2_000_000:
        dense = np.sum(orc.to_dense(a) * orc.to_dense(b))
        assert np.allclose(float(val), dense, rtol=1e-10, atol=1e-10 * orc.norm(a) * orc.norm(b))


@pytest.mark.parametrize("path", golden_files("right_orth"))
def test_right_orth_matches_reference(path):
    z = np.load(path)
    cores = orc.as_cores3(load_cores(z, "in"))
    d = len(cores)
    dense0 = orc.to_dense(cores)
    orc.right_orth(cores, d - 1)
    ref = orc.as_cores3(load_cores(z, "after_last_"))
    for c, r in zip(cores, ref):
        assert c.shape == r.shape
        assert np.allclose(c, r, rtol=1e-12, atol=1e-12)
    for j in range(d - 2, 0, -1):
        orc.right_orth(cores, j)
    ref = orc.as_cores3(load_cores(z, "after_all_"))
    for c, r in zip(cores, ref):
        assert c.shape == r.shape
        assert np.allclose(c, r, rtol=1e-11, atol=1e-11)
    # reference test_right_orthogonalization (tests/main_test.py:200-224)
    for k in range(1, d):
        m = cores[k].reshape(cores[k].shape[0], -1)
        g = m @ m.T
        nz = np.abs(np.diag(g)) > 0.5  # zero-padded rows stay zero (algs.py:1679-1685)
        assert np.allclose(g[np.ix_(nz, nz)], np.eye(int(nz.sum())), atol=1e-13)
    assert np.allclose(orc.to_dense(cores), dense0, rtol=1e-12, atol=1e-12 * np.abs(dense0).max())


@pytest.mark.parametrize("path", golden_files("round"))
def test_round_matches_reference(path):
    z = np.load(path)
    cores = orc.as_cores3(load_cores(z, "in"))
    dense = orc.to_dense(cores)
    out, delta = orc.svd_round(copy.deepcopy(cores), float(z["eps"]))
    assert orc.ranks_of(out) == list(z["ranks_out"])
    err = np.linalg.norm(orc.to_dense(out) - dense) / np.linalg.norm(dense)
    assert abs(err - float(z["rel_err"])) <= 1e-10
    assert err <= float(z["eps"]) * 1.0000001 + 1e-13
    ref = orc.as_cores3(load_cores(z, "out"))
    for c, r in zip(out, ref):
        assert c.shape == r.shape
    assert np.allclose(orc.to_dense(out), orc.to_dense(ref), rtol=0, atol=1e-12 * np.linalg.norm(dense))
    assert delta > 0


@pytest.mark.parametrize("path", golden_files("gramsvd"))
def test_gramsvd_round_matches_reference(path):
    """oracle.gramsvd_round against what the reference's tt_gramsvd_round returned (make_golden.py)."""
    z = np.load(path)
    cores = orc.as_cores3(load_cores(z, "in"))
    dense = orc.to_dense(cores)
    out, delta = orc.gramsvd_round(copy.deepcopy(cores), float(z["eps"]))
    assert orc.ranks_of(out) == list(z["ranks_out"])
    err = np.linalg.norm(orc.to_dense(out) - dense) / np.linalg.norm(dense)
    # Gram-SVD is accurate to ~sqrt(machine eps) relative to ||X|| at best
    assert abs(err - float(z["rel_err"])) <= 1e-8
    ref = orc.as_cores3(load_cores(z, "out"))
    for c, r in zip(out, ref):
        assert c.shape == r.shape
    assert np.allclose(orc.to_dense(out), orc.to_dense(ref), rtol=0, atol=1e-8 * np.linalg.norm(dense))
    assert delta > 0


def test_eps_to_rank_cases():
    s = np.array([4.0, 2.0, 1.0, 0.5])
    assert orc.eps_to_rank(s, 0.4) == 4          # nothing can go
    assert orc.eps_to_rank(s, 0.5) == 3          # tail {0.5}
    assert orc.eps_to_rank(s, 1.2) == 2          # tail {1, 0.5} = 1.118
    assert orc.eps_to_rank(s, 100.0) == 1        # everything fits: rank clamps to 1


@pytest.mark.parametrize("path", golden_files("delta_svd"))
def test_delta_svd_matches_reference(path):
    z = np.load(path)
    u, s, v, rem, dout = orc.delta_svd(z["mat"], float(z["delta_in"]), bool(z["with_normalizing"]))
    assert len(s) == len(z["s"])
    assert np.allclose(s, z["s"], rtol=1e-12)
    assert abs(rem - float(z["remaining_delta"])) <= 1e-12 * max(1.0, abs(rem))
    if bool(z["with_normalizing"]):
        assert abs(dout - float(z["delta_out"])) <= 1e-12 * abs(dout)
    else:
        assert dout is None
    assert np.allclose((u * s) @ v, (z["u"] * z["s"]) @ z["v"], atol=1e-11)


@pytest.mark.parametrize("path", golden_files("ttsvd"))
def test_ttsvd_matches_reference(path):
    z = np.load(path)
    dense = z["dense"]
    cores, delta = orc.tt_svd(dense, float(z["eps"]))
    assert orc.ranks_of(cores) == list(z["ranks_out"])
    assert abs(delta - float(z["delta"])) <= 1e-14 * delta
    err = np.linalg.norm(orc.to_dense(cores) - dense) / np.linalg.norm(dense)
    assert abs(err - float(z["rel_err"])) <= 1e-10
    ref = orc.as_cores3(load_cores(z, "out"))
    assert np.allclose(orc.to_dense(ref), orc.to_dense(cores), atol=1e-11 * np.linalg.norm(dense))


def test_tt_add_doubles_ranks():
    rng = np.random.default_rng(0)
    x = orc.rand_tt([4, 5, 6, 3], [2, 3, 2], rng)
    y = orc.tt_add(x, x)
    assert orc.ranks_of(y) == [4, 6, 4]
    assert np.allclose(orc.to_dense(y), 2 * orc.to_dense(x))


def test_flop_models_match_baseline_md():
    # cfg2 numbers
    assert orc.inner_flops([32] * 64, [256] * 63, [256] * 63) == 133_152_407_552
    assert 2 * orc.tt_bytes([32] * 64, [256] * 63) == 2_080_636_928
    assert orc.inner_flops([8] * 20, [32] * 19, [32] * 19) == 18_908_160
    f = orc.round_flops([64] * 50, [256] * 49, [64] + [128] * 47 + [64])
    assert 4.5e11 < f < 5.5e11


@pytest.mark.skipif(not refshim.reference_available(), reason="reference not present (GPU box)")
def test_live_reference_side_by_side():
    pt = refshim.load_reference()
    from pytens import Index, TensorNetwork
    from pytens.algs import tt_svd_round

    np.random.seed(7)
    shape = [5, 4, 6, 3, 5, 4, 3, 4, 5, 6, 2]
    idx = [Index(f"x{i}", n) for i, n in enumerate(shape)]
    a = TensorNetwork.rand_tt(idx, [3, 4, 5, 4, 3, 5, 4, 3, 4, 2])
    b = TensorNetwork.rand_tt(idx, [2, 3, 3, 4, 4, 3, 3, 2, 2, 2])
    ca = orc.as_cores3([a.value(i) for i in range(len(shape))])
    cb = orc.as_cores3([b.value(i) for i in range(len(shape))])
    ref = float(a.inner(b))
    assert abs(float(orc.inner(ca, cb)) - ref) <= 1e-12 * abs(ref)
    y = a + a
    cy = orc.as_cores3([np.array(y.value(i)) for i in range(len(shape))])
    out = tt_svd_round(copy.deepcopy(y), 1e-9)
    mine, _ = orc.svd_round(cy, 1e-9)
    assert orc.ranks_of(mine) == [out.value(i).shape[-1] for i in range(len(shape) - 1)]


def test_gramsvd_host_helpers_match_oracle():
    """The host-side pieces of the device Gram-SVD rounding (decimal rounding of sqrt-eigenvalues, rank rule)
    against the oracle restatement -- no GPU involved."""
    import torch

    from tensor_networks_b200.gramsvd import _rounded_sqrt, eps_to_rank

    rng = np.random.default_rng(0)
    for scale in (1.0, 1e-6, 1e8):
        eig = scale * np.concatenate([np.sort(rng.random(12))[::-1], [1e-17, -3e-18, 0.0]])
        e12, em12 = _rounded_sqrt(torch.from_numpy(eig))
        ref = orc.round_sqrt_eigs(eig)
        assert np.array_equal(e12.numpy(), ref)
        inv = np.zeros_like(ref)
        inv[ref != 0] = 1.0 / ref[ref != 0]
        assert np.array_equal(em12.numpy(), inv)
    for _ in range(50):
        s = np.sort(rng.random(int(rng.integers(1, 9))))[::-1] * 10.0 ** rng.integers(-3, 3)
        eps = float(rng.random() * 2.0 * s[0])
        assert eps_to_rank(s, eps) == orc.eps_to_rank(s, eps)
