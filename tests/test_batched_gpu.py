"""GPU parity for the batched small-rank kernels (one CTA per tensor train)."""

import numpy as np
import pytest
import torch

from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _make(batch, shape, ra, rb, seed):
    rng = np.random.default_rng(seed)
    a = [orc.rand_tt(shape, ra, rng) for _ in range(batch)]
    b = [orc.rand_tt(shape, rb, rng) for _ in range(batch)]
    return a, b


@pytest.mark.parametrize(
    "batch,shape,ra,rb",
    [
        (5, [8] * 20, [32] * 19, [32] * 19),  # BASELINE cfg5 item shape
        (37, [3, 5, 2, 7, 4, 6], [4, 9, 13, 7, 3], [5, 2, 32, 8, 6]),  # ragged ranks, odd n
        (300, [11, 9, 10], [17, 31], [32, 1]),  # more items than resident CTAs, n > 8 warps
        (4, [6], [], []),  # d = 1
        (37, [3, 5, 2, 7, 4, 6], [4, 10, 14, 8, 2], [6, 2, 32, 8, 6]),  # even ragged ranks: TMA boxes with zero fill
        (300, [11, 9, 10], [18, 30], [32, 2]),  # TMA path, more items than resident CTAs
        (7, [5, 4], [6], [12]),  # d = 2: first core + plain last core only
        (9, [4] * 6, [2] * 5, [2] * 5),  # rank 2 throughout
        (3, [5] * 5, [1] * 4, [1] * 4),  # rank-1
        (2, [6] * 4, [40, 33, 36], [8, 40, 8]),  # ranks > 32: large-rank fallback per item
    ],
)
def test_inner_batched_vs_oracle(batch, shape, ra, rb):
    from tensor_networks_b200.batch import TensorTrainBatch

    a, b = _make(batch, shape, ra, rb, 42)
    ta, tb = TensorTrainBatch.from_numpy(a), TensorTrainBatch.from_numpy(b)
    got = ta.inner(tb).cpu().numpy()
    ref = np.array([float(orc.inner(x, y)) for x, y in zip(a, b)])
    assert got.shape == (batch,)
    assert np.all(np.abs(got - ref) <= RTOL * np.abs(ref)), np.max(np.abs(got - ref) / np.abs(ref))
    nrm = ta.norm().cpu().numpy()
    ref_n = np.array([orc.norm(x) for x in a])
    assert np.all(np.abs(nrm - ref_n) <= RTOL * ref_n)


def test_inner_batched_tma_matches_cp_async_kernel(monkeypatch):
    """The TMA-staged kernel (batched_tma.cu) and the cp.async kernel (batched.cu) agree to roundoff on the
    cfg5 item shape and on a ragged even-rank shape; TTB_BINNER_TMA=0 selects the latter."""
    from tensor_networks_b200.batch import TensorTrainBatch

    for shape, ra, rb in (([8] * 20, [32] * 19, [32] * 19), ([5, 8, 3, 9], [8, 24, 6], [30, 16, 32])):
        ta = TensorTrainBatch.rand(700, shape, ra, seed=11)
        tb = TensorTrainBatch.rand(700, shape, rb, seed=12)
        monkeypatch.setenv("TTB_BINNER_TMA", "1")
        v_tma = ta.inner(tb)
        monkeypatch.setenv("TTB_BINNER_TMA", "0")
        v_cp = ta.inner(tb)
        monkeypatch.delenv("TTB_BINNER_TMA")
        scale = ta.norm() * tb.norm()
        assert torch.all((v_tma - v_cp).abs() <= 1e-13 * scale), ((v_tma - v_cp).abs() / scale).max()


def test_inner_batched_scatter_single_rank():
    """The fused compute + gather entry point (ttb_inner_batched_scatter_f64) with this process as its only peer:
    same values as the plain call, at the right offset, for the TMA kernel and for the fallback shapes."""
    from tensor_networks_b200.batch import TensorTrainBatch
    from tensor_networks_b200.sharding import PeerGather, inner_sharded

    for shape, ra, rb in (([8] * 20, [32] * 19, [32] * 19), ([5, 8, 3, 9], [7, 24, 5], [30, 15, 32]), ([6] * 4, [40, 33, 36], [8, 40, 8])):
        nb = 50 if max(ra) <= 32 else 3
        ta = TensorTrainBatch.rand(nb, shape, ra, seed=21)
        tb = TensorTrainBatch.rand(nb, shape, rb, seed=22)
        ref = ta.inner(tb)
        pg = PeerGather(nb)
        assert pg.fused and pg.ptrs == [pg.tensor.data_ptr()]
        got = inner_sharded(ta, tb, nb, gather=pg)
        assert got.data_ptr() == pg.tensor.data_ptr()
        assert torch.equal(got, ref)
        # two "peers" (two local arrays) and a non-zero offset
        big1 = torch.full((nb + 7,), -1.0, dtype=torch.float64, device="cuda")
        big2 = torch.full((nb + 7,), -1.0, dtype=torch.float64, device="cuda")
        ta.inner_scatter(tb, [big1.data_ptr(), big2.data_ptr()], 4)
        for big in (big1, big2):
            assert torch.equal(big[4 : 4 + nb], ref)
            assert torch.all(big[:4] == -1.0) and torch.all(big[4 + nb :] == -1.0)


def test_inner_batched_matches_single_path_and_is_deterministic():
    from tensor_networks_b200.batch import TensorTrainBatch

    ta = TensorTrainBatch.rand(64, [8] * 20, [32] * 19, seed=4000)
    tb = TensorTrainBatch.rand(64, [8] * 20, [32] * 19, seed=4001)
    v1 = ta.inner(tb)
    v2 = ta.inner(tb)
    assert torch.equal(v1, v2)  # fixed summation order
    single = torch.stack([ta.item(i).inner_dev(tb.item(i)) for i in range(0, 64, 9)])
    assert torch.all((v1[::9] - single).abs() <= RTOL * single.abs())
    # shards are slices of dimension 0
    s0, s1 = ta.shard(0, 2), ta.shard(1, 2)
    t0, t1 = tb.shard(0, 2), tb.shard(1, 2)
    assert torch.equal(torch.cat([s0.inner(t0), s1.inner(t1)]), v1)


# ----------------------------------------------------------------------------- rounding
def _round_case(batch, shape, xr, eps, seed, mode="double"):
    import copy

    from tensor_networks_b200.batch import TensorTrainBatch

    rng = np.random.default_rng(seed)
    items = []
    for _ in range(batch):
        x = orc.rand_tt(shape, xr, rng)
        if mode == "double":
            y = orc.tt_add(x, x)
        else:
            y = x
            for j in range(1, 4):
                z = orc.rand_tt(shape, [2] * (len(shape) - 1), rng)
                z[0] = z[0] * 10.0 ** (-2 * j)
                y = orc.tt_add(y, z)
        items.append(y)
    tb = TensorTrainBatch.from_numpy(items)
    tb.round(eps)
    torch.cuda.synchronize()
    ranks = tb.item_ranks.cpu().numpy()
    assert int(tb.round_status.sum().item()) == 0
    for i, y in enumerate(items):
        ref, _ = orc.svd_round(copy.deepcopy(y), eps)
        assert list(ranks[i][1:-1]) == orc.ranks_of(ref), (i, list(ranks[i]), orc.ranks_of(ref))
        assert ranks[i][0] == 1 and ranks[i][-1] == 1
        if np.prod(shape) <= 2_000_000:
            dense = orc.to_dense(y)
            got = tb.item(i).dense()
            err = np.linalg.norm(got - dense) / np.linalg.norm(dense)
            err_ref = np.linalg.norm(orc.to_dense(ref) - dense) / np.linalg.norm(dense)
            assert abs(err - err_ref) <= 1e-10, (i, err, err_ref)
    return tb


@pytest.mark.parametrize(
    "batch,shape,xr,eps,mode",
    [
        (9, [4] * 6, [3] * 5, 1e-8, "double"),
        (5, [8, 3, 5, 8, 2], [4, 7, 6, 3], 1e-10, "double"),
        (6, [6] * 6, [5] * 5, 1e-2, "decay"),  # genuine truncation
        (3, [7, 5], [4], 1e-8, "double"),  # d = 2
        (200, [5] * 4, [2] * 3, 1e-6, "double"),  # more items than SMs
        (2, [3, 2, 2, 3], [4, 7, 5], 1e-8, "double"),  # n*b < r (pad branch of the reference)
        (4, [8] * 8, [14] * 7, 1e-8, "double"),  # bonds 28: Cholesky-QR fast path with deflation candidates
        (4, [8] * 6, [12] * 5, 1e-4, "decay"),  # bonds 18, graded: grey-zone pivots -> Householder fallback
        (3, [9, 7, 8, 6, 9], [11, 13, 9, 10], 1e-9, "double"),  # odd tile extents (hlen not a multiple of 4 / 8)
    ],
)
def test_round_batched_vs_oracle(batch, shape, xr, eps, mode):
    _round_case(batch, shape, xr, eps, 77, mode)


def test_round_batched_cfg5_items():
    """Items of BASELINE cfg5: d=20, n=8, X bonds 16 doubled to 32 -> [8, 16, ..., 16, 8]."""
    from tensor_networks_b200.batch import TensorTrainBatch

    B, d, n, r = 24, 20, 8, 16
    x = TensorTrainBatch.rand(B, [n] * d, [r] * (d - 1), seed=4000)
    y = x + x
    assert y.ranks() == [2 * r] * (d - 1)
    ny = y.norm()
    z = y.clone().round(1e-8)
    ranks = z.item_ranks.cpu().numpy()
    expect = [1, 8] + [16] * (d - 3) + [8, 1]
    assert (ranks == np.array(expect)[None, :]).all(), ranks[0]
    for i in (0, 7, 23):
        zi, yi = z.item(i), y.item(i)
        nz = zi.norm()
        assert abs(nz - float(ny[i])) <= 1e-10 * nz
        assert abs(float(zi.inner(yi)) / (nz * nz) - 1.0) < 1e-12
        # against the large-rank single-TT path
        single = yi.clone().round(1e-8)
        assert single.ranks() == zi.ranks()
        for c in zi.cores[:-1]:
            m = c.reshape(-1, c.shape[2])
            assert float((m.T @ m - torch.eye(m.shape[1], device=m.device, dtype=m.dtype)).abs().max()) < 1e-12


def test_round_batched_large_rank_fallback():
    _round_case(2, [6] * 4, [20, 24, 18], 1e-8, 5, "double")


def test_pack_rounded_cores_uniform_layout():
    """All-gather of core results (north_star item 4): the pack kernel turns the per-item compact cores
    of a rounded batch into one zero-padded array per core; every item still is the same tensor."""
    from tensor_networks_b200.batch import TensorTrainBatch
    from tensor_networks_b200.sharding import all_gather_cores, padded_ranks, round_sharded

    rng = np.random.default_rng(17)
    shape = [4, 5, 3, 4]
    items = []
    for i in range(6):  # items of different true ranks, stored with the same capacity ranks (6, 6, 6)
        x = orc.rand_tt(shape, [1 + i % 3, 2 + i % 2, 1 + (i + 1) % 3], rng)
        pad = orc.rand_tt(shape, [6 - c.shape[2] for c in x[:-1]], rng)
        pad[0] = pad[0] * 0.0
        items.append(orc.tt_add(x, pad))
    y = TensorTrainBatch.from_numpy(items)
    table, full = round_sharded(y, 1e-10, 6, gather_cores=True)  # single process: gathers are identities
    ranks = table.cpu().numpy()
    assert sorted(set(ranks[:, 1].tolist())) == [1, 2, 3]
    rcap = padded_ranks(table)
    assert full.bond_ranks() == rcap and full.batch == 6 and full.item_ranks is None
    for i in range(6):
        want = orc.to_dense(items[i])
        got = full.item(i).dense()
        assert np.linalg.norm(got - want) <= 1e-9 * np.linalg.norm(want)
    same = all_gather_cores(y, 6, table)
    assert all(torch.equal(a, b) for a, b in zip(same.cores, full.cores))
    # the fused path (pack kernel stores into every rank's arena) with this process as its only peer
    from tensor_networks_b200.sharding import PeerGather, gathered_cores_numel

    arena = PeerGather(gathered_cores_numel(y, 6, table) + 3)
    fused = all_gather_cores(y, 6, table, arena=arena)
    assert all(torch.equal(a, b) for a, b in zip(fused.cores, full.cores))
    assert fused.cores[0].data_ptr() == arena.tensor.data_ptr()
