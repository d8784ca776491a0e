"""CPU tests of the multi-GPU host logic: world_size-2 gloo, oracle as the local compute."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from tensor_networks_b200.sharding import shard_range

    for batch in (0, 1, 7, 8, 8192, 8191):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (l0, h0), (l1, h1) in zip(spans, spans[1:]):
                assert h0 == l1
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, batch, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import tt_oracle as orc
    from tensor_networks_b200.sharding import all_gather_items, shard_range

    rng = np.random.default_rng(1234)  # same stream on every rank -> same global batch
    shape, ra, rb = [3, 4, 2, 3], [2, 3, 2], [3, 2, 2]
    items = [(orc.rand_tt(shape, ra, rng), orc.rand_tt(shape, rb, rng)) for _ in range(batch)]
    lo, hi = shard_range(batch, rank, world)
    local = torch.tensor([float(orc.inner(a, b)) for a, b in items[lo:hi]], dtype=torch.float64)
    full = all_gather_items(local, batch)
    ref = torch.tensor([float(orc.inner(a, b)) for a, b in items], dtype=torch.float64)
    ok_scalar = bool(torch.equal(full, ref))
    # rank tables (int64, 2-d) gather the same way
    ranks_local = torch.arange(lo, hi, dtype=torch.int64)[:, None] * torch.ones(1, 5, dtype=torch.int64)
    ranks_full = all_gather_items(ranks_local, batch)
    ok_ranks = bool(torch.equal(ranks_full[:, 0], torch.arange(batch, dtype=torch.int64)))
    # uniformly padded cores (what all_gather_cores moves after the pack kernel) and the padded ranks
    from tensor_networks_b200.sharding import all_gather_padded_cores, padded_ranks

    table = torch.tensor([[1, 1 + (i % 3), 2 + (i % 2), 1] for i in range(batch)], dtype=torch.int64)
    rcap = padded_ranks(table)
    ok_cap = rcap == [1, min(3, batch), 3 if batch > 1 else 2, 1]
    gen = torch.Generator().manual_seed(7)
    full_cores = [torch.randn((batch, rcap[k], 4, rcap[k + 1]), dtype=torch.float64, generator=gen) for k in range(3)]
    got = all_gather_padded_cores([c[lo:hi].clone() for c in full_cores], batch)
    ok_cores = all(torch.equal(g, f) for g, f in zip(got, full_cores))
    # PeerGather without peer-mappable memory (CPU / gloo): not fused, and inner_sharded takes the collective path
    # into the gather's own buffer
    from tensor_networks_b200.sharding import PeerGather, inner_sharded

    class _Local:  # stands in for a TensorTrainBatch shard: the oracle is the local compute
        batch = hi - lo

        def inner(self, other):
            return local

    pg = PeerGather(batch, device="cpu")
    ok_pg = (not pg.fused) and pg.why_not is not None
    got_pg = inner_sharded(_Local(), _Local(), batch, gather=pg)
    ok_pg = ok_pg and bool(torch.equal(got_pg, ref))
    if batch % world == 0:  # even shards: the collective writes straight into the gather's buffer
        ok_pg = ok_pg and got_pg.data_ptr() == pg.tensor.data_ptr()
    q.put((rank, ok_scalar and ok_ranks and ok_cap and ok_cores and ok_pg))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [7, 8])
def test_all_gather_items_gloo_world2(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + batch
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
