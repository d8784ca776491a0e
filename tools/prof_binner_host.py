"""Host-side cost of one TensorTrainBatch.inner call (tiny batch: the kernel takes ~0.1 ms, the rest is Python / ctypes /
tensor-map encoding), and the enqueue-only cost (no synchronisation)."""
import sys, time, torch
sys.path.insert(0, '.')
from tensor_networks_b200.batch import TensorTrainBatch
d, n, r = 20, 8, 32
a = TensorTrainBatch.rand(148, [n] * d, [r] * (d - 1), seed=1)
b = TensorTrainBatch.rand(148, [n] * d, [r] * (d - 1), seed=2)
out = torch.empty(148, dtype=torch.float64, device='cuda')
for _ in range(5): a.inner(b, out=out)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(200): a.inner(b, out=out)
t_enq = (time.perf_counter() - t) / 200
torch.cuda.synchronize()
t_all = (time.perf_counter() - t) / 200
print(f"enqueue {1e6 * t_enq:.1f} us per call, with completion {1e6 * t_all:.1f} us per call")
