"""CPU tests of the host-side graph logic of the network API (no arithmetic, no GPU): the
structural halves of TensorNetwork.svd / merge (compute_data=False, what the structure search
uses to enumerate candidates: pytens/search/state.py:194-225), attach(), rand_tree's draw order,
the TT constructors -- against fixtures generated from the reference."""

import numpy as np

from conftest import golden_files
from oracle import netio


def _classes():
    from tensor_networks_b200.algs import Index, Tensor, TensorNetwork

    return TensorNetwork, Tensor, Index


def _load(z, prefix):
    return netio.unpack(z, prefix, *_classes())


def test_rand_tree_same_draws_as_reference():
    from tensor_networks_b200.algs import Index, rand_tree

    z = np.load(golden_files("tree_split")[0])
    np.random.seed(100)  # tests/main_test.py:481-486
    tree = rand_tree([Index("x", 5), Index("u", 10), Index("v", 20)], [1, 2, 3, 4, 5])
    assert netio.structure(tree) == netio.meta_structure(z, "in_")
    ref = _load(z, "in_")
    for n in tree.network.nodes:
        assert np.array_equal(tree.value(n), ref.value(n))


def test_svd_and_merge_structure_only():
    from tensor_networks_b200.algs import SVDConfig

    z = np.load(golden_files("tree_split")[5])
    assert not bool(z["compute_data"])
    tree = _load(z, "in_")
    node = int(z["node"])
    (u, s, v), rem = tree.svd(node, [int(i) for i in z["lefts"]], SVDConfig(compute_data=False))
    assert [str(u), str(s), str(v)] == list(z["names"])
    assert rem == float(z["remaining_delta"])
    assert netio.structure(tree) == netio.meta_structure(z, "out_")
    z = np.load(golden_files("tree_merge")[0])
    tree = _load(z, "in_")
    tree.merge(2, 3, compute_data=False)
    assert netio.structure(tree) == netio.meta_structure(z, "merged_nodata_")


def test_attach_structure_and_accessors():
    z = np.load(golden_files("attach")[0])
    a, b = _load(z, "a_"), _load(z, "b_")
    att = a.attach(b)
    assert netio.structure(att) == netio.meta_structure(z, "att_")
    assert [i.name for i in att.free_indices()] == list(z["inner_ab_names"])
    # attach copies: the operands are untouched
    att.network.nodes["G0"]["tensor"].value[:] = 0.0
    assert np.abs(a.value(0)).max() > 0
    assert a.dim() == 3 and a.ranks() == [3, 2] and a.shape() == [4, 5, 6]
    assert a.cost() == 4 * 3 + 3 * 5 * 2 + 2 * 6
    assert a.fresh_index() == "s_0" and a.fresh_node() == "n0"
    assert [i.name for i in a.get_contraction_index(0, 1)] == ["r1"]
    assert a.node_by_free_index("y") == 1


def test_tree_round_fixture_structures_are_trees():
    import networkx as nx

    for path in golden_files("tree_round"):
        z = np.load(path)
        for prefix in ("a_", "comb_", "round_"):
            tn = _load(z, prefix)
            assert nx.is_tree(tn.network)


def test_constructors_host():
    from tensor_networks_b200.algs import Index, TensorNetwork, tt_rank1, tt_separable, ttop_rank1, vector

    idx = [Index("a", 3), Index("b", 4), Index("c", 5)]
    vals = [np.arange(3.0), np.arange(4.0), np.arange(5.0)]
    r1 = tt_rank1(idx, vals)
    assert [r1.value(k).shape for k in range(3)] == [(3, 1), (1, 4, 1), (1, 5)]
    sep = tt_separable(idx, vals)
    dense = np.einsum("ar,rbs,sc->abc", sep.value(0), sep.value(1), sep.value(2))
    want = vals[0][:, None, None] + vals[1][None, :, None] + vals[2][None, None, :]
    assert np.allclose(dense, want)
    op = ttop_rank1(idx, [Index("ap", 3), Index("bp", 4), Index("cp", 5)], [np.eye(3), np.eye(4), np.eye(5)], "A")
    assert [op.value(k).shape for k in range(3)] == [(3, 3, 1), (1, 4, 4, 1), (1, 5, 5)]
    assert [i.name for i in op.node_tensor(1).indices] == ["A_r1", "bp", "b", "A_r2"]
    v = vector("w", idx[0], vals[0])
    assert v.free_indices() == [idx[0]]
    np.random.seed(0)
    t = TensorNetwork.rand_tucker(idx, 2)
    assert sorted(map(str, t.network.nodes)) == ["G0", "G1", "G2", "root"]


def test_tt_compatibility_check_host_logic():
    """Which pairs of networks take the fused TT sweep in TensorNetwork.inner (pytens/algs.py:585-587 semantics are
    kept for everything else): same free indices node by node, TT core shapes, chain of bonds."""
    from tensor_networks_b200.algs import Index, Tensor, TensorNetwork, _tt_signature

    idx = [Index(f"x{i}", 4) for i in range(5)]
    a = TensorNetwork.rand_tt(idx, [2, 3, 3, 2])
    b = TensorNetwork.rand_tt(idx, [3, 1, 2, 4])
    assert _tt_signature(a) == idx and a._tt_compatible(b) and b._tt_compatible(a) and a._tt_compatible(a)
    # a different free index at one node, a different number of nodes, a renamed copy
    other = list(idx)
    other[2] = Index("y", 4)
    assert not a._tt_compatible(TensorNetwork.rand_tt(other, [2, 2, 2, 2]))
    assert not a._tt_compatible(TensorNetwork.rand_tt(idx[:4], [2, 2, 2]))
    # same free indices in a different node order: not the same train
    perm = [idx[1], idx[0]] + idx[2:]
    assert not a._tt_compatible(TensorNetwork.rand_tt(perm, [2, 2, 2, 2]))
    # not a chain in the reference's layout: string node names, a star, a 3-d first core
    named = TensorNetwork()
    named.add_node("a", Tensor(np.zeros((4, 2)), [idx[0], Index("r1", 2)]))
    named.add_node("b", Tensor(np.zeros((2, 4)), [Index("r1", 2), idx[1]]))
    named.add_edge("a", "b")
    assert _tt_signature(named) is None
    odd = TensorNetwork()
    odd.add_node(0, Tensor(np.zeros((1, 4, 2)), [Index("r0", 1), idx[0], Index("r1", 2)]))
    odd.add_node(1, Tensor(np.zeros((2, 4)), [Index("r1", 2), idx[1]]))
    odd.add_edge(0, 1)
    assert _tt_signature(odd) is None
    # a bond that does not link neighbours
    broken = TensorNetwork()
    broken.add_node(0, Tensor(np.zeros((4, 2)), [idx[0], Index("r1", 2)]))
    broken.add_node(1, Tensor(np.zeros((2, 4, 2)), [Index("q", 2), idx[1], Index("r2", 2)]))
    broken.add_node(2, Tensor(np.zeros((2, 4)), [Index("r2", 2), idx[2]]))
    assert _tt_signature(broken) is None
    # a repeated free index is not a TT over d distinct modes
    rep = TensorNetwork.rand_tt([idx[0], idx[0], idx[1]], [2, 2])
    assert not rep._tt_compatible(rep)
