"""GPU tests of the pytens-compatible surface; they read like the reference's own tests
(tests/main_test.py: test_inner :119-126, test_right_orthogonalization :200-224,
test_rounding :226-243, test_scale :340-349) but assert the ranks for real."""

import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tts():
    from tensor_networks_b200.algs import Index, TensorNetwork

    np.random.seed(4)
    x, u, v = Index("x", 5), Index("u", 10), Index("v", 20)
    tt_ranks, tt_ranks2 = [2, 2], [3, 4]
    TT = TensorNetwork.rand_tt([x, u, v], tt_ranks)
    TT2 = TensorNetwork.rand_tt([x, u, v], tt_ranks2)
    return TT, TT2, tt_ranks, tt_ranks2


def test_inner(tts):
    TT, TT2, _, _ = tts
    inner_val = TT.inner(TT2)
    assert isinstance(inner_val, np.ndarray) and inner_val.shape == ()
    out1 = TT.contract().value
    out2 = TT2.contract().value
    assert np.allclose(inner_val, np.sum(out1 * out2), atol=1e-12, rtol=1e-12)
    assert np.isclose(TT.norm(), np.sqrt(np.sum(out1 * out1)), rtol=1e-12)
    assert isinstance(TT.norm(), float)


def test_right_orthogonalization(tts):
    from tensor_networks_b200.algs import tt_right_orth

    TT, _, tt_ranks, _ = tts
    TTc = copy.deepcopy(TT)
    arr1 = TTc.contract().value
    out = tt_right_orth(TTc, 2)
    assert out is TTc
    node = TTc.value(2)
    assert np.allclose(np.dot(node, node.T), np.eye(tt_ranks[1]), atol=1e-14, rtol=1e-14)
    assert np.allclose(arr1, TTc.contract().value, atol=1e-13, rtol=1e-13)
    TTc = tt_right_orth(TTc, 1)
    node = TTc.value(1)
    check = sum(np.dot(node[:, ii, :], node[:, ii, :].T) for ii in range(node.shape[1]))
    assert np.allclose(check, np.eye(tt_ranks[0]), atol=1e-14, rtol=1e-14)
    assert np.allclose(arr1, TTc.contract().value, atol=1e-13, rtol=1e-13)


def test_rounding(tts):
    from tensor_networks_b200.algs import tt_svd_round

    TT, _, tt_ranks, _ = tts
    TTadd = TT + TT
    assert TTadd.ranks() == [4, 4]
    ttadd = TTadd.contract().value
    out = tt_svd_round(TTadd, 1e-5)
    assert out is TTadd  # mutates and returns the same object
    assert TTadd.ranks() == tt_ranks  # the reference's assertTrue(a, b) never checked this
    assert np.allclose(TTadd.contract().value, ttadd, atol=1e-12, rtol=1e-12)
    # indices were resized like Tensor.update_val_size does (pytens/algs.py:70-78)
    assert [i.size for i in TTadd.node_tensor(1).indices] == [2, 10, 2]


def test_scale_and_round_kwargs(tts):
    from tensor_networks_b200 import algs

    TT, _, _, _ = tts
    n0 = TT.norm()
    TT.scale(3.0)
    assert np.isclose(TT.norm(), 3.0 * n0, rtol=1e-13)
    big = TT + TT
    algs.round(big, 1e-10, max_rank=1)
    assert big.ranks() == [1, 1]


def test_tt_svd_and_errors():
    from tensor_networks_b200 import algs

    rng = np.random.default_rng(0)
    dense = np.einsum("ia,ajb,bk->ijk", rng.standard_normal((6, 3)), rng.standard_normal((3, 7, 2)),
                      rng.standard_normal((2, 5)))
    tn = algs.tt_svd(dense, 1e-10)
    assert tn.ranks() == [3, 2]
    assert np.allclose(tn.contract().value, dense, atol=1e-12 * np.abs(dense).max() * 10)
    assert [i.name for i in tn.free_indices()] == ["x0", "x1", "x2"]
    # non-TT networks take the node-level route (attach + contract on the device), like the reference
    other = algs.TensorNetwork()
    other.add_node("a", algs.Tensor(np.ones((2, 2)), [algs.Index("i", 2), algs.Index("j", 2)]))
    assert abs(other.norm() - 2.0) < 1e-14
    # the TT sweeps proper still need a TT-shaped network
    with pytest.raises(NotImplementedError):
        algs.tt_svd_round(other, 1e-3)
    # tree rounding of the TT-shaped result (pytens/algs.py:763-827) works on it as on any tree
    before = tn.contract().value
    free = [i.name for i in tn.free_indices()]
    tn.round(0, 1e-8)
    t = tn.contract()
    have = [i.name for i in t.indices]
    assert np.allclose(np.transpose(t.value, [have.index(n) for n in free]), before, atol=1e-7)


def test_dropin_inner_small_host_trains_packed_path():
    """algs.TensorNetwork.inner on small numpy trains (the points of examples/inner_product_scaling.py) goes through the
    packed upload (all cores in one pinned buffer, one copy): values as the oracle's, odd core sizes and rank-1 ends
    included, operands untouched."""
    from oracle import tt_oracle as orc
    from tensor_networks_b200 import algs
    from tensor_networks_b200.types import Index

    np.random.seed(4)
    for d, n, r in ((20, 20, 10), (12, 5, 5), (6, 7, 3), (2, 9, 4)):
        idx = [Index(f"x{i}", n) for i in range(d)]
        a = algs.TensorNetwork.rand_tt(idx, [r] * (d - 1))
        b = algs.TensorNetwork.rand_tt(idx, [r + 1] * (d - 1))
        ca = [a.network.nodes[k]["tensor"].value.copy() for k in range(d)]
        cb = [b.network.nodes[k]["tensor"].value.copy() for k in range(d)]
        ref = float(orc.inner(orc.as_cores3(ca), orc.as_cores3(cb)))
        got = a.inner(b)
        assert isinstance(got, np.ndarray) and got.shape == () and got.dtype == np.float64
        assert abs(float(got) - ref) <= 1e-12 * abs(ref)
        assert abs(float(b.inner(a)) - ref) <= 1e-12 * abs(ref)
        assert all(np.array_equal(a.network.nodes[k]["tensor"].value, ca[k]) for k in range(d))
    # a network whose free indices differ is not TT-compatible: the reference's attach() + contract() route
    idx2 = [Index(f"x{i}", 5) for i in range(3)]
    idx3 = [Index("x0", 5), Index("x1", 5), Index("y", 5)]
    p, q = algs.TensorNetwork.rand_tt(idx2, [2, 2]), algs.TensorNetwork.rand_tt(idx3, [2, 2])
    assert not p._tt_compatible(q) and p._tt_compatible(p)
