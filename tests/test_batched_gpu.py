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
