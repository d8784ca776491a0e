"""GPU parity for the TT inner product / norm (north_star gate: 1e-12 relative)."""

import numpy as np
import pytest
import torch

from conftest import golden_files, load_cores
from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-12  # north_star: inner products within 1e-12 relative of the reference fp64 path


def _tt(cores):
    from tensor_networks_b200 import TensorTrain

    return TensorTrain.from_cores(cores)


@pytest.mark.parametrize("path", golden_files("inner"))
def test_inner_golden(path):
    z = np.load(path)
    a, b = _tt(load_cores(z, "a")), _tt(load_cores(z, "b"))
    val = a.inner(b)
    assert isinstance(val, np.ndarray) and val.shape == () and val.dtype == np.float64
    ref = float(z["inner"])
    assert abs(float(val) - ref) <= RTOL * abs(ref)
    assert abs(float(b.inner(a)) - ref) <= RTOL * abs(ref)
    assert abs(a.norm() - float(z["norm_a"])) <= RTOL * float(z["norm_a"])
    assert abs(b.norm() - float(z["norm_b"])) <= RTOL * float(z["norm_b"])


@pytest.mark.parametrize(
    "shape,ra,rb",
    [
        ([8] * 20, [32] * 19, [32] * 19),  # one item of BASELINE cfg5
        ([32] * 10, [128] * 9, [128] * 9),
        ([7, 3, 9, 4, 6], [5, 11, 3, 8], [6, 2, 13, 4]),  # ragged, odd ranks
        ([20] * 6, [40] * 5, [17] * 5),  # unequal ranks -> both contraction orders
        ([20] * 6, [9] * 5, [64] * 5),
        ([5], [], []),  # d = 1
        ([4, 6], [3], [5]),
    ],
)
def test_inner_vs_oracle(shape, ra, rb):
    rng = np.random.default_rng(123)
    a = orc.rand_tt(shape, ra, rng)
    b = orc.rand_tt(shape, rb, rng)
    ref = float(orc.inner(a, b))
    got = float(_tt(a).inner(_tt(b)))
    # random independent TTs have a tiny cosine; the gate is relative to the value itself
    assert abs(got - ref) <= RTOL * abs(ref), (got, ref)
    na = _tt(a).norm()
    assert abs(na - orc.norm(a)) <= RTOL * na


def test_dense_contraction():
    rng = np.random.default_rng(5)
    a = orc.rand_tt([5, 4, 6, 3, 7], [3, 8, 5, 2], rng, scaled=False)
    got = _tt(a).dense()
    ref = orc.to_dense(a)
    assert got.shape == ref.shape
    assert np.allclose(got, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
    b = orc.rand_tt([5, 4, 6, 3, 7], [2, 2, 2, 2], rng, scaled=False)
    # reference test_inner (tests/main_test.py:119-126): inner == sum(dense * dense)
    val = float(_tt(a).inner(_tt(b)))
    assert np.isclose(val, np.sum(ref * orc.to_dense(b)), rtol=1e-12)


def test_inner_properties_full_size():
    """BASELINE cfg2 shape (d=64, n=32, r=256): size-independent properties."""
    from tensor_networks_b200 import TensorTrain

    d, n, r = 64, 32, 256
    a = TensorTrain.rand([n] * d, [r] * (d - 1), seed=1001)
    b = TensorTrain.rand([n] * d, [r] * (d - 1), seed=1002)
    ab = float(a.inner(b))
    ba = float(b.inner(a))
    assert np.isfinite(ab) and ab != 0.0
    assert abs(ab - ba) <= 1e-11 * abs(ab)  # symmetry (different contraction operands)
    # homogeneity: scaling one core scales the result exactly by a power of two
    a.scale(4.0)
    assert float(a.inner(b)) == 4.0 * ab
    a.scale(0.25)
    # ||a||^2 > 0 and Cauchy-Schwarz
    na, nb = a.norm(), b.norm()
    assert na > 0 and nb > 0 and abs(ab) <= na * nb
    # linearity in the first argument through the block-diagonal sum: <a + a, b> = 2 <a, b>
    s = TensorTrain.rand([n] * 8, [16] * 7, seed=7)
    t = TensorTrain.rand([n] * 8, [24] * 7, seed=8)
    st = float(s.inner(t))
    assert abs(float((s + s).inner(t)) - 2 * st) <= 1e-12 * abs(st)


def test_inner_quarter_size_vs_oracle():
    """d=16 slice of cfg2 (n=32, r=256) against the numpy oracle."""
    from tensor_networks_b200 import TensorTrain

    d, n, r = 16, 32, 256
    a = TensorTrain.rand([n] * d, [r] * (d - 1), seed=11)
    b = TensorTrain.rand([n] * d, [r] * (d - 1), seed=12)
    ref = float(orc.inner(a.to_cores(), b.to_cores()))
    got = float(a.inner(b))
    assert abs(got - ref) <= RTOL * abs(ref)


def test_errors():
    from tensor_networks_b200 import TensorTrain

    a = TensorTrain.rand([4, 5, 6], [2, 3], seed=1)
    b = TensorTrain.rand([4, 5, 7], [2, 3], seed=2)
    with pytest.raises(AssertionError):
        a.inner(b)


@pytest.mark.parametrize(
    "d,n,r",
    [
        # BASELINE configs[0] = examples/inner_product_scaling.py (the reference's own CPU sweep): rank scaling at
        # n = 20, d = 20 (the two largest ranks on fewer cores so that the numpy oracle finishes in seconds) ...
        (20, 20, 10), (20, 20, 20), (20, 20, 40), (20, 20, 80), (20, 20, 160), (8, 20, 320), (4, 20, 640),
        # ... mode-size scaling at r = 20, d = 20 ...
        (20, 5, 20), (20, 160, 20), (20, 2560, 20),
        # ... and dimension scaling at r = 5, n = 5 (cores scaled: the unscaled d = 640 product overflows fp64)
        (5, 5, 5), (80, 5, 5), (640, 5, 5),
    ],
)
def test_inner_reference_scaling_sweep(d, n, r):
    rng = np.random.default_rng(4)  # the example seeds numpy with 4
    a = orc.rand_tt([n] * d, [r] * (d - 1), rng)
    b = orc.rand_tt([n] * d, [r] * (d - 1), rng)
    ref = float(orc.inner(a, b))
    got = float(_tt(a).inner(_tt(b)))
    assert np.isfinite(ref) and ref != 0.0
    assert abs(got - ref) <= RTOL * abs(ref), (got, ref)
