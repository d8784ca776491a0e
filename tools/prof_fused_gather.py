"""torchrun script: cost of the pieces of the fused gather of batched inner products (kernel with peer stores,
signal barrier, NCCL all-gather) on a strong-scaled batch of 8192 pairs."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
from tensor_networks_b200.batch import TensorTrainBatch
from tensor_networks_b200.sharding import PeerGather, all_gather_items, inner_sharded, shard_range
batch, d, n, r = 8192, 20, 8, 32
lo, hi = shard_range(batch, rank, world)
a = TensorTrainBatch.rand(hi - lo, [n] * d, [r] * (d - 1), seed=1 + rank)
b = TensorTrainBatch.rand(hi - lo, [n] * d, [r] * (d - 1), seed=101 + rank)
pg = PeerGather(batch)
vals = torch.empty(batch, dtype=torch.float64, device="cuda")
def timed(f, reps=20):
    for _ in range(3): f()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
res = {
    "kernel_local": timed(lambda: a.inner(b)),
    "kernel_peer_stores": timed(lambda: a.inner_scatter(b, pg.ptrs, lo)) if pg.fused else None,
    "signal_barrier": timed(lambda: pg.barrier()) if pg.handle is not None else None,
    "nccl_all_gather_8B_per_item": timed(lambda: all_gather_items(vals[lo:hi], batch, out=vals)),
    "fused_step": timed(lambda: inner_sharded(a, b, batch, gather=pg)) if pg.fused else None,
    "nccl_step": timed(lambda: all_gather_items(a.inner(b), batch, out=vals)),
}
if rank == 0: print(world, {k: (round(v, 4) if v is not None else None) for k, v in res.items()}, pg.why_not)
dist.destroy_process_group()
