"""Generate tests/golden/*.npz from the reference itself -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Every fixture stores the seeded INPUT cores and what the unmodified reference
functions returned for them (through the import shim in `oracle/refshim.py`):
`TensorNetwork.inner` / `norm` (pytens/algs.py:585-594), `tt_right_orth`
(:1654-1704), `tt_svd_round` (:1841-1903), `tt_gramsvd_round` (:1771-1838), `delta_svd` (pytens/utils.py:19-100)
and the TT-SVD composition `TensorNetwork.svd` + `merge` (:633-702, :735-761).
The fixtures are what pins `oracle/tt_oracle.py` (tests/test_oracle.py) and the
CUDA path (tests/test_*gpu*.py) on the GPU box, where the reference is absent.
"""

from __future__ import annotations

import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

pt = refshim.load_reference()
from pytens import Index, SVDConfig, Tensor, TensorNetwork  # noqa: E402
from pytens.algs import tt_gramsvd_round, tt_right_orth, tt_svd_round  # noqa: E402
from pytens.utils import delta_svd  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def ref_cores(tn):
    d = len(tn.network.nodes)
    return [np.array(tn.value(i)) for i in range(d)]


def pack(prefix, cores):
    return {f"{prefix}{k}": c for k, c in enumerate(cores)}


def scaled_rand_tt(shape, ranks, scale=True):
    idx = [Index(f"x{i}", n) for i, n in enumerate(shape)]
    tt = TensorNetwork.rand_tt(idx, list(ranks))
    if scale:
        d = len(shape)
        r = list(ranks) + [1]
        for k in range(d):
            t = tt.node_tensor(k)
            t.value = t.value / np.sqrt(shape[k] * r[k])
    return tt


def gen_inner():
    cases = [
        # (seed, shape, ranksA, ranksB, scaled)
        (4, [5, 10, 20], [2, 2], [3, 4], False),  # tests/main_test.py:119-126 shapes
        (11, [4, 3, 5, 2, 6], [3, 4, 2, 5], [2, 2, 3, 3], False),
        (12, [6] * 8, [7] * 7, [6] * 7, False),  # SURVEY 3.1 probe shape
        (13, [8] * 12, [9] * 11, [5] * 11, True),  # d >= 10: beyond 26 letters
        (14, [3, 7], [4], [6], False),  # d = 2
        (15, [5] * 6, [1] * 5, [3] * 5, False),  # rank-1 operand
        (16, [2] * 16, [4] * 15, [4] * 15, True),
    ]
    for i, (seed, shape, ra, rb, scaled) in enumerate(cases):
        np.random.seed(seed)
        a = scaled_rand_tt(shape, ra, scaled)
        b = scaled_rand_tt(shape, rb, scaled)
        val = a.inner(b)
        assert isinstance(val, np.ndarray) and val.shape == ()
        d = dict(
            shape=np.array(shape),
            inner=np.array(val),
            norm_a=np.array(a.norm()),
            norm_b=np.array(b.norm()),
            inner_aa=np.array(a.inner(a)),
        )
        d.update(pack("a", ref_cores(a)))
        d.update(pack("b", ref_cores(b)))
        np.savez(os.path.join(OUT, f"inner_{i}.npz"), **d)
        print("inner", i, float(val))


def gen_right_orth():
    cases = [
        (21, [5, 10, 20], [2, 2]),  # tests/main_test.py:200-224
        (22, [4, 3, 5, 6], [6, 9, 4]),
        (23, [3, 2, 2, 3], [4, 7, 5]),  # n*b < r on core 2 -> zero-pad path
        (24, [6, 6, 3], [5, 8]),  # last core r > n -> shrink path
    ]
    for i, (seed, shape, ranks) in enumerate(cases):
        np.random.seed(seed)
        tt = scaled_rand_tt(shape, ranks, False)
        d = dict(shape=np.array(shape))
        d.update(pack("in", ref_cores(tt)))
        dn = len(shape)
        tt = tt_right_orth(tt, dn - 1)
        d.update(pack("after_last_", ref_cores(tt)))
        for j in range(dn - 2, 0, -1):
            tt = tt_right_orth(tt, j)
        d.update(pack("after_all_", ref_cores(tt)))
        np.savez(os.path.join(OUT, f"right_orth_{i}.npz"), **d)
        print("right_orth", i, [c.shape for c in ref_cores(tt)])


def gen_round():
    cases = [
        # (seed, shape, ranks of X, eps, mode) ; Y = X + X like tests/main_test.py:228
        (31, [5, 10, 20], [2, 2], 1e-5, "double"),
        (32, [6] * 6, [3, 5, 5, 4, 2], 1e-8, "double"),
        (33, [4] * 10, [4] * 9, 1e-10, "double"),
        (34, [8] * 5, [6, 6, 6, 6], 1e-3, "noise"),  # genuine truncation
        (35, [3, 2, 2, 3, 4], [4, 7, 5, 3], 1e-8, "double"),  # pad path
        (36, [7, 5], [4], 1e-8, "double"),  # d = 2
        (37, [8] * 8, [8] * 7, 1e-2, "decay"),  # decaying spectrum, heavy truncation
    ]
    for i, (seed, shape, ranks, eps, mode) in enumerate(cases):
        np.random.seed(seed)
        x = scaled_rand_tt(shape, ranks, True)
        if mode == "double":
            y = x + x
        elif mode == "noise":
            z = scaled_rand_tt(shape, ranks, True)
            z.scale(1e-5)
            y = x + z
        else:
            # sum of TTs with geometrically decaying weights
            y = x
            for j in range(1, 4):
                z = scaled_rand_tt(shape, [2] * (len(shape) - 1), True)
                z.scale(10.0 ** (-2 * j))
                y = y + z
        dense = y.contract().value
        d = dict(shape=np.array(shape), eps=np.array(eps))
        d.update(pack("in", ref_cores(y)))
        out = tt_svd_round(copy.deepcopy(y), eps)
        oc = ref_cores(out)
        ranks_out = [c.shape[-1] for c in oc[:-1]]
        dense_out = out.contract().value
        err = np.linalg.norm(dense_out - dense) / np.linalg.norm(dense)
        d.update(pack("out", oc))
        d.update(
            ranks_out=np.array(ranks_out),
            rel_err=np.array(err),
            norm_in=np.array(np.linalg.norm(dense)),
        )
        np.savez(os.path.join(OUT, f"round_{i}.npz"), **d)
        print("round", i, [c.shape[-1] for c in ref_cores(y)[:-1]], "->", ranks_out, err)


def gen_delta_svd():
    rng = np.random.default_rng(41)
    cases = []
    # tall (m > 10 n), square-ish, wide, rank-deficient, normalising
    a = rng.standard_normal((200, 8)) @ np.diag(10.0 ** -np.arange(8)) @ rng.standard_normal((8, 8))
    cases.append((a, 1e-4, False))
    b = rng.standard_normal((30, 20))
    cases.append((b, 2.0, False))
    c = rng.standard_normal((6, 50))
    cases.append((c, 1e-3, True))
    dmat = rng.standard_normal((40, 5)) @ rng.standard_normal((5, 12))
    cases.append((dmat, 1e-9, True))
    e = rng.standard_normal((12, 12))
    cases.append((e, 1e3, False))  # everything truncated -> rank clamps to 1
    for i, (mat, delta, wn) in enumerate(cases):
        t = delta_svd(mat, delta, wn)
        np.savez(
            os.path.join(OUT, f"delta_svd_{i}.npz"),
            mat=mat,
            delta_in=np.array(delta),
            with_normalizing=np.array(wn),
            u=t.u,
            s=t.s,
            v=t.v,
            remaining_delta=np.array(t.remaining_delta),
            delta_out=np.array(np.nan if t.delta is None else t.delta),
        )
        print("delta_svd", i, mat.shape, "rank", len(t.s))


def ref_tt_svd(dense, eps):
    """TT-SVD composed from the reference's own svd/merge."""
    shape = dense.shape
    d = len(shape)
    delta = eps / np.sqrt(d - 1) * np.linalg.norm(dense.ravel())
    net = TensorNetwork()
    net.add_node("G", Tensor(dense.copy(), [Index(f"x{i}", n) for i, n in enumerate(shape)]))
    node = "G"
    order = []
    lefts = [0]
    for _ in range(d - 1):
        (u, s, v), _ = net.svd(node, lefts, SVDConfig(delta=delta, with_orthonormal=False))
        net.merge(v, s)
        order.append(u)
        node = v
        nd = len(net.node_tensor(v).indices)
        lefts = [nd - 1, 0]
    order.append(node)
    cores = []
    for k, name in enumerate(order):
        val = net.value(name)
        if k == 0:
            cores.append(val.reshape(1, shape[0], -1))
        elif k == d - 1:
            # indices are (x_{d-1}, bond): bond last after merge(v, s)
            cores.append(np.ascontiguousarray(val.T).reshape(-1, shape[k], 1))
        else:
            cores.append(val)
    return cores, delta, net


def gen_ttsvd():
    cases = [
        (51, [6, 6, 6, 6, 6], [3, 4, 4, 3], 1e-10, 0.0),  # SURVEY 3.3 probe
        (52, [4, 5, 3, 6], [2, 3, 2], 1e-8, 1e-12),
        (53, [3, 3, 3, 3, 3, 3, 3], [3, 5, 7, 7, 5, 3], 1e-3, 1e-4),
        (54, [10, 12], [4], 1e-10, 0.0),
    ]
    for i, (seed, shape, ranks, eps, noise) in enumerate(cases):
        np.random.seed(seed)
        x = scaled_rand_tt(shape, ranks, True)
        dense = x.contract().value
        if noise > 0:
            dense = dense + noise * np.linalg.norm(dense) / np.sqrt(dense.size) * np.random.randn(*dense.shape)
        cores, delta, net = ref_tt_svd(dense, eps)
        approx = net.contract()
        # bring the reference result into x0..x{d-1} order
        names = [ix.name for ix in approx.indices]
        perm = [names.index(f"x{k}") for k in range(len(shape))]
        rec = np.transpose(approx.value, perm)
        err = np.linalg.norm(rec - dense) / np.linalg.norm(dense)
        ranks_out = [c.shape[2] for c in cores[:-1]]
        d = dict(
            dense=dense,
            eps=np.array(eps),
            delta=np.array(delta),
            ranks_out=np.array(ranks_out),
            rel_err=np.array(err),
        )
        d.update(pack("out", cores))
        np.savez(os.path.join(OUT, f"ttsvd_{i}.npz"), **d)
        print("ttsvd", i, shape, "->", ranks_out, err)


def gen_gramsvd():
    cases = [
        # (seed, shape, ranks of X, eps, mode); first case = tests/main_test.py:245-262
        (61, [5, 10, 20], [2, 2], 1e-5, "double"),
        (62, [6] * 6, [3, 5, 5, 4, 2], 1e-5, "double"),
        (63, [8] * 5, [6, 6, 6, 6], 1e-3, "noise"),
        (64, [8] * 8, [8] * 7, 1e-2, "decay"),
        (65, [7, 5], [4], 1e-6, "double"),
        (66, [4] * 10, [4] * 9, 1e-6, "double"),
    ]
    for i, (seed, shape, ranks, eps, mode) in enumerate(cases):
        np.random.seed(seed)
        x = scaled_rand_tt(shape, ranks, True)
        if mode == "double":
            y = x + x
        elif mode == "noise":
            z = scaled_rand_tt(shape, ranks, True)
            z.scale(1e-5)
            y = x + z
        else:
            y = x
            for j in range(1, 4):
                z = scaled_rand_tt(shape, [2] * (len(shape) - 1), True)
                z.scale(10.0 ** (-2 * j))
                y = y + z
        dense = y.contract().value
        d = dict(shape=np.array(shape), eps=np.array(eps))
        d.update(pack("in", ref_cores(y)))
        out = tt_gramsvd_round(copy.deepcopy(y), eps)
        oc = ref_cores(out)
        ranks_out = [c.shape[-1] for c in oc[:-1]]
        dense_out = out.contract().value
        err = np.linalg.norm(dense_out - dense) / np.linalg.norm(dense)
        d.update(pack("out", oc))
        d.update(ranks_out=np.array(ranks_out), rel_err=np.array(err), norm_in=np.array(np.linalg.norm(dense)))
        np.savez(os.path.join(OUT, f"gramsvd_{i}.npz"), **d)
        print("gramsvd", i, [c.shape[-1] for c in ref_cores(y)[:-1]], "->", ranks_out, err)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gramsvd":  # add the Gram-SVD fixtures without touching the others
        gen_gramsvd()
        sys.exit(0)
    gen_inner()
    gen_right_orth()
    gen_round()
    gen_delta_svd()
    gen_ttsvd()
    gen_gramsvd()
    print("fixtures written to", OUT)
