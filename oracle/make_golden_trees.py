"""Generate the tree / network / TT-operator fixtures (tests/golden/tree_*.npz, tensor_*.npz,
attach_*.npz, ttop_*.npz, randround_*.npz) from the reference itself -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python oracle/make_golden_trees.py

Covers the rows the first set of fixtures (oracle/make_golden.py) does not: `Tensor.svd / qr /
contract / permute / mult / block_diagonal` (pytens/algs.py:143-344), `TensorNetwork.svd / qr /
merge / round / orthonormalize / attach / + / *` (:521-572, :633-955, :1310-1380) on the shapes of
the reference's own tests (tests/main_test.py:456-514, :642-987), `tt_sum`, `ttop_*`, `gmres`
(:2383-2793) and the randomised rounding (:2133-2380, seeded).  Every fixture stores the inputs and
what the unmodified reference returned.
"""

from __future__ import annotations

import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import netio  # noqa: E402
import refshim  # noqa: E402

pt = refshim.load_reference()
from pytens import Index, SVDConfig, Tensor, TensorNetwork  # noqa: E402
from pytens.algs import (  # noqa: E402
    gmres, rand_tree, tt_rand_precond_svd_round, tt_randomized_round, tt_sum, tt_sum_randomized_round,
    ttop_apply, ttop_rank1, ttop_rank2, ttop_sum, ttop_sum_apply,
)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def save(name, **arrays):
    np.savez(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name)


def names_of(indices):
    return np.array([str(i.name) for i in indices])


def dense_of(tn, order):
    return netio.dense_in_order(tn, order)


# --------------------------------------------------------------------------- Tensor-level
def gen_tensor_ops():
    rng = np.random.default_rng(71)
    # svd / qr: (shape, lefts, delta)
    cases = [
        ((6, 7, 8), [0, 2], 1e-5),
        ((4, 5, 6, 3), [3, 1], 0.5),      # real truncation, permuted lefts
        ((12, 3), [0], 1e-9),
        ((3, 40), [0], 1e-9),             # wide
        ((5, 4, 3, 2), [2], 2.0),
        ((30, 4, 2), [0, 1], 1e-8),        # tall (m > 10 n): QR-first branch of delta_svd
    ]
    for i, (shape, lefts, delta) in enumerate(cases):
        val = rng.standard_normal(shape)
        if i == 1:  # graded so that delta = 0.5 truncates something but not everything
            val = val * np.linspace(1.0, 0.01, shape[-1])
        t = Tensor(val.copy(), [Index(f"i{k}", n) for k, n in enumerate(shape)])
        [u, s, v], rem = t.svd(lefts, delta)
        q, r = t.qr(lefts)
        save(f"tensor_split_{i}", value=val, lefts=np.array(lefts), delta=np.array(delta),
             u=u.value, s=s.value, v=v.value, remaining_delta=np.array(rem),
             u_names=names_of(u.indices), v_names=names_of(v.indices), s_names=names_of(s.indices),
             q=q.value, r=r.value, q_names=names_of(q.indices), r_names=names_of(r.indices))
    # contract / permute / mult / block_diagonal / concat_fill
    a = Tensor(rng.standard_normal((4, 5, 6)), [Index("a", 4), Index("b", 5), Index("c", 6)])
    b = Tensor(rng.standard_normal((6, 3, 4)), [Index("c", 6), Index("d", 3), Index("a", 4)])
    c = a.contract(b)
    b2 = Tensor(rng.standard_normal((5, 7)), [Index("b", 5), Index("e", 7)])
    c2 = a.contract(b2)
    b3 = Tensor(rng.standard_normal((2, 3)), [Index("p", 2), Index("q", 3)])
    c3 = a.contract(b3)  # no common index: outer product
    p = a.permute([2, 0, 1])
    x = Tensor(rng.standard_normal((2, 5, 3)), [Index("r0", 2), Index("x", 5), Index("r1", 3)])
    y = Tensor(rng.standard_normal((4, 5, 2)), [Index("s0", 4), Index("x", 5), Index("s1", 2)])
    m = x.mult(y, [Index("x", 5)])
    bd = x.block_diagonal(y, [Index("x", 5)])
    cf = x.concat_fill(y, [Index("x", 5)])
    save("tensor_pair_0", a=a.value, b=b.value, c=c.value, c_names=names_of(c.indices),
         b2=b2.value, c2=c2.value, c2_names=names_of(c2.indices), b3=b3.value, c3=c3.value,
         c3_names=names_of(c3.indices), p=p.value, p_names=names_of(p.indices),
         x=x.value, y=y.value, mult=m.value, mult_names=names_of(m.indices),
         mult_sizes=np.array([i.size for i in m.indices]),
         bd=bd.value, bd_names=names_of(bd.indices), cf=cf.value, cf_names=names_of(cf.indices))


# --------------------------------------------------------------------------- network-level
def ref_test_tree():
    np.random.seed(100)
    return rand_tree([Index("x", 5), Index("u", 10), Index("v", 20)], [1, 2, 3, 4, 5])


def gen_tree_split():
    # tests/main_test.py:488-514 (test_tree_split, test_tree_split_free) + variations
    cases = [
        ("ref", 4, [0, 2], SVDConfig()),
        ("ref", 3, [0, 1], SVDConfig()),
        ("ref", 4, [0, 2], SVDConfig(delta=1e-8, with_orthonormal=False)),
        ("big", None, None, SVDConfig(delta=900.0, with_orthonormal=True)),
        ("big", None, None, SVDConfig(delta=8.0, with_orthonormal=False)),
        ("ref", 4, [0, 2], SVDConfig(compute_data=False)),
    ]
    for i, (kind, node, lefts, cfg) in enumerate(cases):
        if kind == "ref":
            tree = ref_test_tree()
        else:
            np.random.seed(7 + i)
            tree = rand_tree([Index("a", 8), Index("b", 9), Index("c", 10), Index("d", 11)], [3, 4, 5, 6, 4])
            node = max(tree.network.nodes, key=lambda n: len(tree.node_tensor(n).indices))
            nd = len(tree.node_tensor(node).indices)
            lefts = [nd - 1, 0] if nd > 2 else [0]
        free = [ix.name for ix in tree.free_indices()]
        d = netio.pack(tree, "in_")
        d.update(node=np.array(netio._enc(node)[1]), lefts=np.array(lefts), delta=np.array(cfg.delta),
                 with_orthonormal=np.array(cfg.with_orthonormal), compute_data=np.array(cfg.compute_data),
                 free=np.array(free), dense=dense_of(tree, free))
        work = copy.deepcopy(tree)
        (u, s, v), rem = work.svd(node, lefts, cfg)
        d.update(netio.pack(work, "out_", with_values=cfg.compute_data))
        d.update(names=np.array([str(u), str(s), str(v)]), remaining_delta=np.array(rem))
        if cfg.compute_data:
            d.update(dense_out=dense_of(work, free))
        save(f"tree_split_{i}", **d)
    # merge (tests/main_test.py:516-530) and qr on the same tree
    tree = ref_test_tree()
    free = [ix.name for ix in tree.free_indices()]
    d = netio.pack(tree, "in_")
    work = copy.deepcopy(tree)
    work.merge(2, 3)
    d.update(netio.pack(work, "merged_"))
    work2 = copy.deepcopy(tree)
    qn, rn = work2.qr(4, [0, 2])
    d.update(netio.pack(work2, "qr_"))
    d.update(qr_names=np.array([str(qn), str(rn)]), free=np.array(free), dense=dense_of(tree, free))
    work3 = copy.deepcopy(tree)
    work3.merge(2, 3, compute_data=False)
    d.update(netio.pack(work3, "merged_nodata_", with_values=False))
    save("tree_merge_0", **d)


def nets_add1():
    x = Tensor(np.random.randn(2, 5, 6), [Index("a", 2), Index("i", 5), Index("j", 6)])
    u = Tensor(np.random.randn(2, 7), [Index("a", 2), Index("k", 7)])
    n1 = TensorNetwork(); n1.add_node("x", x); n1.add_node("u", u); n1.add_edge("x", "u")
    y = Tensor(np.random.randn(3, 5, 6), [Index("b", 3), Index("i", 5), Index("j", 6)])
    v = Tensor(np.random.randn(3, 7), [Index("b", 3), Index("k", 7)])
    n2 = TensorNetwork(); n2.add_node("y", y); n2.add_node("v", v); n2.add_edge("y", "v")
    return n1, n2, "x"


def nets_star():
    def star(center, cname, cinds, leaves):
        net = TensorNetwork()
        net.add_node(cname, Tensor(np.random.randn(*[s for _, s in cinds]), [Index(n, s) for n, s in cinds]))
        for lname, (bn, bs), (fn, fs) in leaves:
            net.add_node(lname, Tensor(np.random.randn(bs, fs), [Index(bn, bs), Index(fn, fs)]))
            net.add_edge(cname, lname)
        return net
    n1 = star(None, "x", [("a", 1), ("b", 2), ("c", 3), ("d", 4)],
              [("u1", ("a", 1), ("i", 5)), ("u2", ("b", 2), ("j", 6)), ("u3", ("c", 3), ("k", 7)),
               ("u4", ("d", 4), ("l", 8))])
    n2 = star(None, "y", [("e", 2), ("f", 3), ("g", 4), ("h", 5)],
              [("v1", ("e", 2), ("i", 5)), ("v2", ("f", 3), ("j", 6)), ("v3", ("g", 4), ("k", 7)),
               ("v4", ("h", 5), ("l", 8))])
    return n1, n2, "x"


def nets_deep(names1=("x", "u1", "u2", "u3", "u4"), names2=("y", "v1", "v2", "v3", "v4"), free=("k", "m", "l")):
    fk, fm, fl = free
    a = TensorNetwork()
    a.add_node(names1[0], Tensor(np.random.randn(5, 6, 2, 5), [Index("i", 5), Index("j", 6), Index("a", 2), Index("b", 5)]))
    a.add_node(names1[1], Tensor(np.random.randn(2, 7), [Index("d", 2), Index(fk, 7)]))
    a.add_node(names1[2], Tensor(np.random.randn(5, 8), [Index("b", 5), Index(fm, 8)]))
    a.add_node(names1[3], Tensor(np.random.randn(2, 3, 2), [Index("a", 2), Index("c", 3), Index("d", 2)]))
    a.add_node(names1[4], Tensor(np.random.randn(3, 9), [Index("c", 3), Index(fl, 9)]))
    a.add_edge(names1[0], names1[3]); a.add_edge(names1[0], names1[2])
    a.add_edge(names1[3], names1[1]); a.add_edge(names1[3], names1[4])
    b = TensorNetwork()
    b.add_node(names2[0], Tensor(np.random.randn(5, 6, 1, 2), [Index("i", 5), Index("j", 6), Index("aa", 1), Index("bb", 2)]))
    b.add_node(names2[1], Tensor(np.random.randn(3, 7), [Index("dd", 3), Index(fk, 7)]))
    b.add_node(names2[2], Tensor(np.random.randn(2, 8), [Index("bb", 2), Index(fm, 8)]))
    b.add_node(names2[3], Tensor(np.random.randn(1, 2, 3), [Index("aa", 1), Index("cc", 2), Index("dd", 3)]))
    b.add_node(names2[4], Tensor(np.random.randn(2, 9), [Index("cc", 2), Index(fl, 9)]))
    b.add_edge(names2[0], names2[2]); b.add_edge(names2[0], names2[3])
    b.add_edge(names2[3], names2[1]); b.add_edge(names2[3], names2[4])
    return a, b, names1[0]


def gen_tree_round():
    """tests/main_test.py:456-477 (test_optimize) and :642-987 (test_add1-4, test_mul1-3): same
    topologies, node names, index names and bond sizes; the free mode sizes are 5..9 instead of
    13..17 so that the stored dense tensors stay small."""
    cases = []
    np.random.seed(4)
    tt = TensorNetwork.rand_tt([Index("x", 5), Index("y", 10), Index("z", 20)], [2, 2])
    cases.append(("add", tt, copy.deepcopy(tt), 0, None, 1e-5))          # test_optimize: absolute delta 1e-5
    np.random.seed(101); n1, n2, root = nets_add1(); cases.append(("add", n1, n2, root, 1e-10, None))
    np.random.seed(102); n1, n2, root = nets_star(); cases.append(("add", n1, n2, root, 1e-10, None))
    np.random.seed(103); n1, n2, root = nets_deep(); cases.append(("add", n1, n2, root, 1e-10, None))
    np.random.seed(104); n1, _n2, root = nets_deep(); cases.append(("add", n1, copy.deepcopy(n1), root, 1e-10, None))
    np.random.seed(105); n1, n2, root = nets_add1(); cases.append(("mul", n1, n2, root, 1e-10, None))
    np.random.seed(106); n1, n2, root = nets_star(); cases.append(("mul", n1, n2, root, 1e-10, None))
    np.random.seed(107)
    n1, n2, root = nets_deep(("u0", "u1", "u2", "u3", "u4"), ("v0", "v1", "v2", "v3", "v4"), ("k", "l", "m"))
    cases.append(("mul", n1, n2, root, 1e-10, None))
    np.random.seed(108); n1, n2, root = nets_deep(); cases.append(("sub", n1, n2, root, 1e-6, None))
    for i, (op, n1, n2, root, rel, absd) in enumerate(cases):
        free = [ix.name for ix in n1.free_indices()]
        t1, t2 = dense_of(n1, free), dense_of(n2, free)
        comb = n1 + n2 if op == "add" else (n1 * n2 if op == "mul" else n1 - n2)
        want = t1 + t2 if op == "add" else (t1 * t2 if op == "mul" else t1 - t2)
        d = netio.pack(n1, "a_"); d.update(netio.pack(n2, "b_")); d.update(netio.pack(comb, "comb_"))
        nrm = comb.norm()
        delta = absd if absd is not None else nrm * rel
        orth = copy.deepcopy(comb)
        orth_root = orth.orthonormalize(root)
        d.update(netio.pack(orth, "orth_", with_values=False))
        work = copy.deepcopy(comb)
        ret, rem = work.round(root, delta)
        d.update(netio.pack(work, "round_"))
        d.update(op=np.array(op), root=np.array(netio._enc(root)[1]), free=np.array(free), dense=want,
                 norm=np.array(nrm), delta=np.array(delta), orth_root=np.array(str(orth_root)),
                 round_ret=np.array(str(ret)), remaining_delta=np.array(rem),
                 dense_round=dense_of(work, free), cost_in=np.array(comb.cost()), cost_out=np.array(work.cost()))
        save(f"tree_round_{i}", **d)


def gen_attach():
    """inner() of networks whose free indices are only partly shared, and of non-TT networks
    (attach + contract, pytens/algs.py:521-587)."""
    np.random.seed(81)
    x, y, z, w = Index("x", 4), Index("y", 5), Index("z", 6), Index("w", 3)
    a = TensorNetwork.rand_tt([x, y, z], [3, 2])
    b = TensorNetwork.rand_tt([x, y, w], [2, 4])
    r = a.attach(b).contract()
    np.random.seed(82)
    n1, n2, _ = nets_add1()
    r2 = n1.attach(n2).contract()
    att = a.attach(b)
    d = netio.pack(a, "a_"); d.update(netio.pack(b, "b_")); d.update(netio.pack(n1, "c_")); d.update(netio.pack(n2, "d_"))
    d.update(netio.pack(att, "att_", with_values=False))
    d.update(inner_ab=a.inner(b), inner_ab_names=names_of(r.indices), inner_cd=np.array(n1.inner(n2)),
             inner_cd_names=names_of(r2.indices), norm_c=np.array(n1.norm()))
    save("attach_0", **d)


# --------------------------------------------------------------------------- TT operators / GMRES
def gen_ttops():
    np.random.seed(91)
    x, y, z = Index("x", 5), Index("y", 4), Index("z", 3)
    xo, yo, zo = Index("xp", 5), Index("yp", 4), Index("zp", 3)
    mats1 = [np.random.randn(5, 5), np.random.randn(4, 4), np.random.randn(3, 3)]
    mats2 = [np.random.randn(5, 5), np.random.randn(4, 4), np.random.randn(3, 3)]
    mats3 = [np.random.randn(5, 5), np.random.randn(4, 4), np.random.randn(3, 3)]
    tt = TensorNetwork.rand_tt([x, y, z], [3, 2])
    tt2 = TensorNetwork.rand_tt([x, y, z], [2, 4])
    tt3 = TensorNetwork.rand_tt([x, y, z], [1, 2])
    op1 = ttop_rank1([x, y, z], [xo, yo, zo], mats1, "A")
    op2 = ttop_rank2([x, y, z], [xo, yo, zo], mats1, mats2, "B")
    op3 = ttop_sum([x, y, z], [xo, yo, zo], [mats1, mats2, mats3], "C")
    d = {}
    for name, net in (("tt", tt), ("tt2", tt2), ("tt3", tt3), ("op1", op1), ("op2", op2), ("op3", op3),
                      ("app1", ttop_apply(op1, tt)), ("app2", ttop_apply(op2, tt)), ("app3", ttop_apply(op3, tt)),
                      ("sum", tt_sum([tt, tt2, tt3]))):
        d.update(netio.pack(net, name + "_"))
    funcs = [[(lambda v, m=m: np.einsum("ij,jk->ik", m, v)) if k == 0 else
              ((lambda v, m=m: np.einsum("jk,mkp->mjp", m, v)) if k == 1 else
               (lambda v, m=m: np.einsum("jk,mk->mj", m, v)))
              for k, m in enumerate(ms)] for ms in (mats1, mats2, mats3)]
    sa = ttop_sum_apply(tt, [x, y, z], [xo, yo, zo], funcs, "D")
    d.update(netio.pack(sa, "sumapply_"))
    for j, ms in enumerate((mats1, mats2, mats3)):
        for k, m in enumerate(ms):
            d[f"mat_{j}_{k}"] = m
    save("ttop_0", **d)
    # tests/main_test.py:428-448 (test_gmres)
    np.random.seed(92)
    xi, yi, zi = Index("x", 10), Index("y", 5), Index("z", 3)
    A = np.random.randn(10, 10) + 4.0 * np.eye(10)
    ttop = ttop_rank1([xi, yi, zi], [Index("xp", 10), Index("yp", 5), Index("zp", 3)], [A, np.eye(5), np.eye(3)], "A")
    rhs = TensorNetwork.rand_tt([xi, yi, zi], [3, 2])
    x0 = TensorNetwork.rand_tt([xi, yi, zi], [3, 2])
    sol, resid = gmres(lambda t: ttop_apply(ttop, t), rhs, x0, 1e-5, 1e-10, maxiter=30)
    g = netio.pack(rhs, "rhs_"); g.update(netio.pack(x0, "x0_")); g.update(netio.pack(sol, "sol_"))
    g.update(A=A, resid=np.array(resid), dense_sol=dense_of(sol, ["x", "y", "z"]))
    save("gmres_0", **g)


# --------------------------------------------------------------------------- randomised rounding
def gen_randround():
    cases = [
        (201, [6, 7, 8, 5], [3, 4, 3], [6, 8, 6], "double"),
        (202, [5] * 6, [2, 3, 3, 3, 2], [4, 6, 6, 6, 4], "double"),
        (203, [8] * 5, [5] * 4, [5] * 4, "single"),     # target = true rank: exact
        (204, [6, 7, 8, 5], [3, 4, 3], [2, 3, 2], "single"),   # target below the rank: lossy
    ]
    for i, (seed, shape, ranks, target, mode) in enumerate(cases):
        np.random.seed(seed)
        idx = [Index(f"x{k}", n) for k, n in enumerate(shape)]
        x = TensorNetwork.rand_tt(idx, list(ranks))
        y = x + x if mode == "double" else x
        names = [ix.name for ix in idx]
        d = netio.pack(y, "in_")
        d.update(target=np.array(target), dense=dense_of(y, names), seed_call=np.array(seed + 1000))
        np.random.seed(seed + 1000)
        out = tt_randomized_round(copy.deepcopy(y), list(target))
        d.update(netio.pack(out, "out_"))
        d.update(dense_out=dense_of(out, names))
        np.random.seed(seed + 2000)
        pre = tt_rand_precond_svd_round(copy.deepcopy(y), 1e-8, list(target))
        d.update(netio.pack(pre, "pre_"))
        d.update(dense_pre=dense_of(pre, names))
        # sum of three TTs without forming the sum
        np.random.seed(seed + 3000)
        parts = [TensorNetwork.rand_tt(idx, list(ranks)) for _ in range(3)]
        total = sum(dense_of(p, names) for p in parts)
        for j, p in enumerate(parts):
            d.update(netio.pack(p, f"part{j}_"))
        sum_target = [3 * r for r in ranks]
        np.random.seed(seed + 4000)
        so = tt_sum_randomized_round([copy.deepcopy(p) for p in parts], list(sum_target))
        d.update(netio.pack(so, "sumout_"))
        d.update(sum_target=np.array(sum_target), dense_sum=total, dense_sumout=dense_of(so, names))
        save(f"randround_{i}", **d)


if __name__ == "__main__":
    only = sys.argv[1:] or ["tensor", "split", "round", "attach", "ttops", "randround"]
    if "tensor" in only:
        gen_tensor_ops()
    if "split" in only:
        gen_tree_split()
    if "round" in only:
        gen_tree_round()
    if "attach" in only:
        gen_attach()
    if "ttops" in only:
        gen_ttops()
    if "randround" in only:
        gen_randround()
    print("fixtures written to", OUT)
