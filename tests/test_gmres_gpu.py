"""TT-GMRES on the device: the heaviest caller of inner / norm / tt_svd_round in the reference
(pytens/algs.py:2701-2793).  Mirrors tests/main_test.py:428-448 (residual < 1e-5) and additionally
checks the solution against a dense solve and the operator application against einsum
(tests/main_test.py:352-426)."""

import numpy as np
import pytest

from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def test_ttop_apply_matches_dense():
    from tensor_networks_b200 import TensorTrain
    from tensor_networks_b200.solvers import TTOperator, ttop_apply

    rng = np.random.default_rng(0)
    shape_in, shape_out, r_op, r_tt = [4, 3, 5], [6, 2, 5], [1, 3, 2, 1], [4, 3]
    cores = [rng.standard_normal((r_op[k], shape_out[k], shape_in[k], r_op[k + 1])) for k in range(3)]
    op = TTOperator([torch.from_numpy(c).cuda() for c in cores])
    x = orc.rand_tt(shape_in, r_tt, rng)
    y = ttop_apply(op, TensorTrain.from_cores(x))
    assert y.ranks() == [r_tt[0] * r_op[1], r_tt[1] * r_op[2]]
    a_dense = op.dense()  # (o1, i1, o2, i2, o3, i3)
    should_be = np.einsum("ijklmn,jln->ikm", a_dense, orc.to_dense(x))
    assert np.allclose(y.dense(), should_be, atol=1e-12, rtol=1e-12)


def test_gmres_reference_case():
    """Same setting as the reference's test_gmres: A (x) I (x) I on a 10 x 5 x 3 TT, resid < 1e-5."""
    from tensor_networks_b200 import TensorTrain
    from tensor_networks_b200.solvers import gmres, ttop_apply, ttop_rank1

    rng = np.random.default_rng(4)
    A = rng.standard_normal((10, 10))
    op_tt = ttop_rank1([A, np.eye(5), np.eye(3)])
    b = orc.rand_tt([10, 5, 3], [3, 2], rng)
    x0 = orc.rand_tt([10, 5, 3], [3, 2], rng)
    op = lambda t: ttop_apply(op_tt, t)  # noqa: E731
    x, resid = gmres(op, TensorTrain.from_cores(b), TensorTrain.from_cores(x0), 1e-5, 1e-10, maxiter=30)
    assert resid < 1e-5
    x_true = np.einsum("ij,jkl->ikl", np.linalg.inv(A), orc.to_dense(b))
    assert np.linalg.norm(x.dense() - x_true) <= 1e-5 * max(1.0, np.linalg.norm(x_true)) * np.linalg.cond(A)


def test_gmres_rank2_operator():
    """Sum of two Kronecker terms (operator rank 2) on a 12^4 grid: ranks grow and are rounded back."""
    from tensor_networks_b200 import TensorTrain
    from tensor_networks_b200.solvers import TTOperator, gmres, ttop_apply

    rng = np.random.default_rng(9)
    n, d = 12, 4
    mats1 = [np.eye(n) * 2.0 + 0.1 * rng.standard_normal((n, n)) for _ in range(d)]
    mats2 = [0.2 * rng.standard_normal((n, n)) for _ in range(d)]
    cores = []
    for k in range(d):
        rl, rr = (1 if k == 0 else 2), (1 if k == d - 1 else 2)
        c = np.zeros((rl, n, n, rr))
        c[0, :, :, 0] = mats1[k]
        c[rl - 1, :, :, rr - 1] = c[rl - 1, :, :, rr - 1] + mats2[k] if (rl == 1 or rr == 1) else mats2[k]
        cores.append(c)
    op_tt = TTOperator([torch.from_numpy(c).cuda() for c in cores])
    a_dense = op_tt.dense().transpose(0, 2, 4, 6, 1, 3, 5, 7).reshape(n**d, n**d)
    b = orc.rand_tt([n] * d, [3] * (d - 1), rng)
    x0 = orc.rand_tt([n] * d, [2] * (d - 1), rng)
    op = lambda t: ttop_apply(op_tt, t)  # noqa: E731
    x, resid = gmres(op, TensorTrain.from_cores(b), TensorTrain.from_cores(x0), 1e-16, 1e-12, maxiter=40)
    bn = np.linalg.norm(orc.to_dense(b))
    assert resid <= 1e-6 * bn
    x_true = np.linalg.solve(a_dense, orc.to_dense(b).reshape(-1)).reshape([n] * d)
    assert np.linalg.norm(x.dense() - x_true) <= 1e-5 * np.linalg.norm(x_true)
