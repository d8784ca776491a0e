"""GPU parity for the node-level and tree-level network operations (SURVEY 8(a) rows a2, a8-a10,
8(f) row 1) against fixtures generated from the reference by oracle/make_golden_trees.py:
Tensor.svd / qr / contract / permute / mult / block_diagonal (pytens/algs.py:143-344),
TensorNetwork.svd / qr / merge (:633-761), round / orthonormalize (:763-955), attach / inner on
general networks (:521-587), + - * on trees (:1310-1380).

Structural results (node names, index names and sizes, edges, returned names, cost) must be
IDENTICAL to the reference's; values are compared through gauge-invariant quantities (dense
contraction, singular values, orthonormality) with the tolerances of the reference's own tests or
tighter."""

import copy
import os

import numpy as np
import pytest

from conftest import golden_files
from oracle import netio

pytestmark = pytest.mark.gpu


def _classes():
    from tensor_networks_b200.algs import Index, Tensor, TensorNetwork

    return TensorNetwork, Tensor, Index


def _load(z, prefix):
    return netio.unpack(z, prefix, *_classes())


def _key(z, name):
    v = z[name]
    return int(v) if v.dtype.kind in "iu" else str(v)


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


# --------------------------------------------------------------------------- Tensor level
@pytest.mark.parametrize("path", golden_files("tensor_split"))
def test_tensor_svd_qr(path):
    from tensor_networks_b200.algs import Index, Tensor

    z = np.load(path)
    val = z["value"]
    lefts = [int(i) for i in z["lefts"]]
    t = Tensor(val.copy(), [Index(f"i{k}", n) for k, n in enumerate(val.shape)])
    [u, s, v], rem = t.svd(lefts, float(z["delta"]))
    assert isinstance(u.value, np.ndarray)  # numpy in -> numpy out
    assert [str(i.name) for i in u.indices] == list(z["u_names"])
    assert [str(i.name) for i in s.indices] == list(z["s_names"])
    assert [str(i.name) for i in v.indices] == list(z["v_names"])
    assert u.value.shape == z["u"].shape and v.value.shape == z["v"].shape and s.value.shape == z["s"].shape
    smax = np.diag(z["s"]).max()
    assert np.allclose(np.diag(s.value), np.diag(z["s"]), rtol=1e-10, atol=1e-13 * smax)
    assert np.count_nonzero(s.value - np.diag(np.diag(s.value))) == 0
    assert abs(rem - float(z["remaining_delta"])) <= 1e-9 * max(abs(float(z["remaining_delta"])), smax * 1e-4)
    rank = s.value.shape[0]
    um = u.value.reshape(-1, rank)
    vm = v.value.reshape(rank, -1)
    assert np.allclose(um.T @ um, np.eye(rank), atol=1e-12)
    assert np.allclose(vm @ vm.T, np.eye(rank), atol=1e-12)
    ref = z["u"].reshape(-1, rank) @ z["s"] @ z["v"].reshape(rank, -1)
    assert _rel(um @ s.value @ vm, ref) < 1e-12
    # qr
    q, r = t.qr(lefts)
    assert [str(i.name) for i in q.indices] == list(z["q_names"])
    assert [str(i.name) for i in r.indices] == list(z["r_names"])
    assert q.value.shape == z["q"].shape and r.value.shape == z["r"].shape
    k = q.value.shape[-1]
    qm, rm = q.value.reshape(-1, k), r.value.reshape(k, -1)
    assert np.allclose(qm.T @ qm, np.eye(k), atol=1e-12)
    assert _rel(qm @ rm, z["q"].reshape(-1, k) @ z["r"].reshape(k, -1)) < 1e-13
    # same factorisation up to the signs of the columns of q: |diag R| agrees, R is upper triangular
    kk = min(rm.shape)
    assert np.allclose(np.abs(np.diag(rm[:, :kk])), np.abs(np.diag(z["r"].reshape(k, -1)[:, :kk])), rtol=1e-9, atol=1e-12)
    assert np.abs(np.tril(rm[:, :kk], -1)).max(initial=0.0) < 1e-12 * np.abs(rm).max()


def test_tensor_pair_ops():
    from tensor_networks_b200.algs import Index, Tensor

    z = np.load(golden_files("tensor_pair")[0])
    a = Tensor(z["a"], [Index("a", 4), Index("b", 5), Index("c", 6)])
    b = Tensor(z["b"], [Index("c", 6), Index("d", 3), Index("a", 4)])
    c = a.contract(b)
    assert [i.name for i in c.indices] == list(z["c_names"]) and _rel(c.value, z["c"]) < 1e-14
    c2 = a.contract(Tensor(z["b2"], [Index("b", 5), Index("e", 7)]))
    assert [i.name for i in c2.indices] == list(z["c2_names"]) and _rel(c2.value, z["c2"]) < 1e-14
    c3 = a.contract(Tensor(z["b3"], [Index("p", 2), Index("q", 3)]))
    assert [i.name for i in c3.indices] == list(z["c3_names"]) and _rel(c3.value, z["c3"]) < 1e-15
    p = a.permute([2, 0, 1])
    assert [i.name for i in p.indices] == list(z["p_names"]) and np.array_equal(p.value, z["p"])
    x = Tensor(z["x"], [Index("r0", 2), Index("x", 5), Index("r1", 3)])
    y = Tensor(z["y"], [Index("s0", 4), Index("x", 5), Index("s1", 2)])
    m = x.mult(y, [Index("x", 5)])
    assert [i.name for i in m.indices] == list(z["mult_names"])
    assert [i.size for i in m.indices] == list(z["mult_sizes"])
    assert _rel(m.value, z["mult"]) < 1e-15
    bd = x.block_diagonal(y, [Index("x", 5)])
    assert [i.name for i in bd.indices] == list(z["bd_names"]) and np.array_equal(bd.value, z["bd"])
    cf = x.concat_fill(y, [Index("x", 5)])
    assert [i.name for i in cf.indices] == list(z["cf_names"]) and np.array_equal(cf.value, z["cf"])


def test_tensor_ops_stay_on_device():
    """Device values in -> device values out, nothing visits the host."""
    import torch

    from tensor_networks_b200.algs import Index, Tensor

    g = torch.Generator(device="cuda").manual_seed(3)
    a = Tensor(torch.randn((6, 5, 4), dtype=torch.float64, device="cuda", generator=g),
               [Index("a", 6), Index("b", 5), Index("c", 4)])
    [u, s, v], _ = a.svd([0, 2], 1e-12)
    assert u.value.is_cuda and s.value.is_cuda and v.value.is_cuda
    rec = u.contract(s).contract(v)  # indices: a, c, b
    back = rec.permute([0, 2, 1])
    assert back.value.is_cuda
    assert _rel(back.value.cpu().numpy(), a.value.cpu().numpy()) < 1e-13
    q, r = a.qr([1])
    assert q.value.is_cuda and _rel(q.contract(r).permute([1, 0, 2]).value.cpu().numpy(), a.value.cpu().numpy()) < 1e-13


# --------------------------------------------------------------------------- splits / merges
@pytest.mark.parametrize("path", golden_files("tree_split"))
def test_network_svd(path):
    """TensorNetwork.svd on the reference's test_tree_split shapes (tests/main_test.py:488-514) and
    variations (truncating delta, no orthonormalisation, compute_data=False)."""
    from tensor_networks_b200.algs import SVDConfig

    z = np.load(path)
    tree = _load(z, "in_")
    cfg = SVDConfig(delta=float(z["delta"]), with_orthonormal=bool(z["with_orthonormal"]),
                    compute_data=bool(z["compute_data"]))
    free = [str(n) for n in z["free"]]
    (u, s, v), rem = tree.svd(_key(z, "node"), [int(i) for i in z["lefts"]], cfg)
    assert [str(u), str(s), str(v)] == list(z["names"])
    assert netio.structure(tree) == netio.meta_structure(z, "out_")
    if not cfg.compute_data:
        assert rem == float(z["remaining_delta"])
        return
    nd = np.linalg.norm(z["dense"])
    got = netio.dense_in_order(tree, free)
    assert abs(rem - float(z["remaining_delta"])) <= 1e-9 * nd
    # same truncated subspace: the split networks represent the same tensor as the reference's
    assert np.linalg.norm(got - z["dense_out"]) <= 1e-11 * nd
    err, err_ref = _rel(got, z["dense"]), _rel(z["dense_out"], z["dense"])
    assert abs(err - err_ref) <= 1e-10
    # reference test: reconstruction within 1e-5 when delta = 1e-5 (test_tree_split)
    if float(z["delta"]) <= 1e-5:
        assert np.allclose(got, z["dense"], atol=1e-5, rtol=1e-5)


def test_network_merge_and_qr():
    z = np.load(golden_files("tree_merge")[0])
    free = [str(n) for n in z["free"]]
    tree = _load(z, "in_")
    tree.merge(2, 3)
    assert netio.structure(tree) == netio.meta_structure(z, "merged_")
    ref = _load(z, "merged_")
    for n in tree.network.nodes:
        assert _rel(tree.value(n), ref.value(n)) < 1e-14
    assert _rel(netio.dense_in_order(tree, free), z["dense"]) < 1e-14
    tree = _load(z, "in_")
    qn, rn = tree.qr(4, [0, 2])
    assert [str(qn), str(rn)] == list(z["qr_names"])
    assert netio.structure(tree) == netio.meta_structure(z, "qr_")
    assert _rel(netio.dense_in_order(tree, free), z["dense"]) < 1e-13
    qt = tree.node_tensor(qn)
    qm = qt.value.reshape(-1, qt.value.shape[-1])
    assert np.allclose(qm.T @ qm, np.eye(qm.shape[1]), atol=1e-12)
    tree = _load(z, "in_")
    tree.merge(2, 3, compute_data=False)
    assert netio.structure(tree) == netio.meta_structure(z, "merged_nodata_")
    tree = _load(z, "in_")
    with pytest.raises(RuntimeError):
        nonadjacent = [(a, b) for a in tree.network.nodes for b in tree.network.nodes
                       if a != b and not tree.network.has_edge(a, b)][0]
        tree.merge(*nonadjacent)


# --------------------------------------------------------------------------- tree rounding
@pytest.mark.parametrize("path", golden_files("tree_round"))
def test_tree_add_mul_round(path):
    """test_optimize / test_add1-4 / test_mul1-3 of the reference (tests/main_test.py:456-477,
    :642-987): network + / * / -, norm, orthonormalize, round."""
    z = np.load(path)
    a, b = _load(z, "a_"), _load(z, "b_")
    op = str(z["op"])
    free = [str(n) for n in z["free"]]
    root = _key(z, "root")
    comb = a + b if op == "add" else (a * b if op == "mul" else a - b)
    assert netio.structure(comb) == netio.meta_structure(z, "comb_")
    ref_comb = _load(z, "comb_")
    for n in comb.network.nodes:
        assert _rel(comb.value(n), ref_comb.value(n)) < 1e-15
    nd = np.linalg.norm(z["dense"])
    assert _rel(netio.dense_in_order(comb, free), z["dense"]) < 1e-13
    nrm = comb.norm()
    assert abs(nrm - float(z["norm"])) <= 1e-12 * float(z["norm"])
    # orthonormalize: same structure and returned root; every other node is an isometry towards the root
    orth = copy.deepcopy(comb)
    orth_root = orth.orthonormalize(root)
    assert str(orth_root) == str(z["orth_root"])
    assert netio.structure(orth) == netio.meta_structure(z, "orth_")
    assert _rel(netio.dense_in_order(orth, free), z["dense"]) < 1e-12
    root_fro = np.linalg.norm(orth.value(orth_root))
    assert abs(root_fro - nd) <= 1e-12 * nd  # all the norm sits in the root
    # round
    work = copy.deepcopy(comb)
    ret, rem = work.round(root, float(z["delta"]))
    assert str(ret) == str(z["round_ret"])
    assert netio.structure(work) == netio.meta_structure(z, "round_")
    assert work.cost() == int(z["cost_out"])
    got = netio.dense_in_order(work, free)
    err, err_ref = _rel(got, z["dense"]), _rel(z["dense_round"], z["dense"])
    assert abs(err - err_ref) <= 1e-10
    assert np.allclose(got, z["dense"], rtol=1e-10, atol=1e-10 * max(1.0, np.abs(z["dense"]).max()))
    assert abs(rem - float(z["remaining_delta"])) <= 1e-6 * float(z["delta"]) + 1e-12 * nd


def test_tree_round_device_resident():
    """The same rounding with every node value on the GPU: results stay there."""
    z = np.load(golden_files("tree_round")[3])
    a, b = _load(z, "a_").to_device(), _load(z, "b_").to_device()
    comb = a + b
    import torch

    assert all(isinstance(comb.value(n), torch.Tensor) and comb.value(n).is_cuda for n in comb.network.nodes)
    comb.round(_key(z, "root"), float(z["delta"]))
    assert all(isinstance(comb.value(n), torch.Tensor) for n in comb.network.nodes)
    assert netio.structure(comb) == netio.meta_structure(z, "round_")
    free = [str(n) for n in z["free"]]
    assert _rel(netio.dense_in_order(comb, free), z["dense"]) < 1e-10


# --------------------------------------------------------------------------- attach / inner
def test_attach_and_general_inner():
    z = np.load(golden_files("attach")[0])
    a, b, c, d = (_load(z, p) for p in ("a_", "b_", "c_", "d_"))
    att = a.attach(b)
    assert netio.structure(att) == netio.meta_structure(z, "att_")
    got = a.inner(b)  # free indices only partly shared: the open ones survive
    assert got.shape == z["inner_ab"].shape
    assert [i.name for i in att.contract().indices] == list(z["inner_ab_names"])
    assert _rel(got, z["inner_ab"]) < 1e-13
    val = c.inner(d)  # non-TT networks, all free indices shared
    assert np.asarray(val).shape == () and abs(float(val) - float(z["inner_cd"])) <= 1e-12 * abs(float(z["inner_cd"]))
    assert abs(c.norm() - float(z["norm_c"])) <= 1e-12 * float(z["norm_c"])
