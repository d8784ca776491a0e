"""GPU parity of the TT-operator / TT-sum / GMRES functions with the reference's signatures
(pytens/algs.py:2383-2793) and of the randomised rounding (row 4: :2133-2380),
against fixtures generated from the reference (oracle/make_golden_trees.py)."""

import copy

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


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def _same_net(got, z, prefix, tol=1e-14):
    assert netio.structure(got) == netio.meta_structure(z, prefix)
    ref = _load(z, prefix)
    for n in got.network.nodes:
        assert got.value(n).shape == ref.value(n).shape
        assert _rel(got.value(n), ref.value(n)) <= tol, n


def test_ttop_builders_apply_and_sum():
    from tensor_networks_b200.algs import (Index, tt_sum, ttop_apply, ttop_rank1, ttop_rank2, ttop_sum,
                                           ttop_sum_apply)

    z = np.load(golden_files("ttop")[0])
    x, y, zz = Index("x", 5), Index("y", 4), Index("z", 3)
    out = [Index("xp", 5), Index("yp", 4), Index("zp", 3)]
    mats = [[z[f"mat_{j}_{k}"] for k in range(3)] for j in range(3)]
    tt, tt2, tt3 = _load(z, "tt_"), _load(z, "tt2_"), _load(z, "tt3_")
    op1 = ttop_rank1([x, y, zz], out, mats[0], "A")
    op2 = ttop_rank2([x, y, zz], out, mats[0], mats[1], "B")
    op3 = ttop_sum([x, y, zz], out, mats, "C")
    _same_net(op1, z, "op1_", 0.0)
    _same_net(op2, z, "op2_", 0.0)
    _same_net(op3, z, "op3_", 0.0)
    _same_net(ttop_apply(op1, tt), z, "app1_")
    _same_net(ttop_apply(op2, tt), z, "app2_")
    _same_net(ttop_apply(op3, tt), z, "app3_")
    _same_net(tt_sum([tt, tt2, tt3]), z, "sum_", 0.0)
    funcs = [[(lambda v, m=m: np.einsum("ij,jk->ik", m, v)) if k == 0 else
              ((lambda v, m=m: np.einsum("jk,mkp->mjp", m, v)) if k == 1 else
               (lambda v, m=m: np.einsum("jk,mk->mj", m, v)))
              for k, m in enumerate(ms)] for ms in mats]
    _same_net(ttop_sum_apply(tt, [x, y, zz], out, funcs, "D"), z, "sumapply_", 1e-15)
    # device-resident operands give device-resident results
    import torch

    app = ttop_apply(copy.deepcopy(op3).to_device(), copy.deepcopy(tt).to_device())
    assert all(isinstance(app.value(n), torch.Tensor) for n in app.network.nodes)
    ref = _load(z, "app3_")
    for n in app.network.nodes:
        assert _rel(app.value(n).cpu().numpy(), ref.value(n)) < 1e-14


def test_gmres_reference_case():
    """tests/main_test.py:428-448: residual below 1e-5; same solution as the reference's."""
    from tensor_networks_b200.algs import Index, gmres, ttop_apply, ttop_rank1

    z = np.load(golden_files("gmres")[0])
    xi, yi, zi = Index("x", 10), Index("y", 5), Index("z", 3)
    ttop = ttop_rank1([xi, yi, zi], [Index("xp", 10), Index("yp", 5), Index("zp", 3)],
                      [z["A"], np.eye(5), np.eye(3)], "A")
    rhs, x0 = _load(z, "rhs_"), _load(z, "x0_")
    sol, resid = gmres(lambda t: ttop_apply(ttop, t), rhs, x0, 1e-5, 1e-10, maxiter=30)
    assert resid < 1e-5
    assert abs(resid - float(z["resid"])) < 1e-6
    got = netio.dense_in_order(sol, ["x", "y", "z"])
    assert _rel(got, z["dense_sol"]) < 1e-5


@pytest.mark.parametrize("path", golden_files("randround"))
def test_randomized_rounding(path):
    """Seeded like the fixture generator: the sketch matrices are the reference's own draws, so the
    rounded trains represent the same tensors (QR bases are unique up to signs)."""
    from tensor_networks_b200.algs import (tt_rand_precond_svd_round, tt_randomized_round,
                                           tt_sum_randomized_round)

    z = np.load(path)
    y = _load(z, "in_")
    d = len(y.network.nodes)
    names = [f"x{k}" for k in range(d)]
    target = [int(t) for t in z["target"]]
    seed = int(z["seed_call"])
    nd = np.linalg.norm(z["dense"])
    np.random.seed(seed)
    out = tt_randomized_round(copy.deepcopy(y), list(target))
    assert netio.structure(out) == netio.meta_structure(z, "out_")
    got = netio.dense_in_order(out, names)
    assert np.linalg.norm(got - z["dense_out"]) <= 1e-9 * nd
    assert abs(_rel(got, z["dense"]) - _rel(z["dense_out"], z["dense"])) <= 1e-9
    for k in range(d - 1):  # cores 0..d-2 come out of a QR: orthonormal columns
        c = out.value(k)
        m = c.reshape(-1, c.shape[-1])
        assert np.allclose(m.T @ m, np.eye(m.shape[1]), atol=1e-12)
    np.random.seed(seed + 1000)
    pre = tt_rand_precond_svd_round(copy.deepcopy(y), 1e-8, list(target))
    assert netio.structure(pre) == netio.meta_structure(z, "pre_")
    assert np.linalg.norm(netio.dense_in_order(pre, names) - z["dense_pre"]) <= 1e-9 * nd
    parts = [_load(z, f"part{j}_") for j in range(3)]
    np.random.seed(seed + 3000)
    so = tt_sum_randomized_round([copy.deepcopy(p) for p in parts], [int(t) for t in z["sum_target"]])
    assert netio.structure(so) == netio.meta_structure(z, "sumout_")
    ns = np.linalg.norm(z["dense_sum"])
    assert np.linalg.norm(netio.dense_in_order(so, names) - z["dense_sumout"]) <= 1e-9 * ns
    assert _rel(netio.dense_in_order(so, names), z["dense_sum"]) < 1e-11
