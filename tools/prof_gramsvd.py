"""Gram-SVD rounding of the generic bond-128 workload (bench_extras.run_gramsvd shape): ms per call and per bond."""
import sys, time, torch
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain
d, n = 20, 64
y = None
for j in range(4):
    t = TensorTrain.rand([n] * d, [32] * (d - 1), seed=5001 + j); t.cores[0].mul_(10.0 ** (-3 * j))
    y = t if y is None else y + t
for _ in range(2):
    z = y.clone().gramsvd_round(1e-5)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    z = y.clone(); torch.cuda.synchronize(); t0 = time.perf_counter(); z.gramsvd_round(1e-5); torch.cuda.synchronize()
    ts.append(1e3 * (time.perf_counter() - t0))
print("gramsvd ms", min(ts), "per bond", min(ts) / (d - 1), "ranks", z.ranks()[:6])
