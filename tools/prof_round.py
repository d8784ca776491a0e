"""Steady-state rounding of the cfg3 structure (shorter train) for ncu: run with
ncu --profile-from-start off ...; only the last round is inside the profiler range."""
import sys, torch, time
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain
d, n, r = int(sys.argv[1]) if len(sys.argv) > 1 else 10, 64, 128
generic = len(sys.argv) > 2 and sys.argv[2] == "generic"
if generic:
    y = None
    for j in range(4):
        t = TensorTrain.rand([n] * d, [32] * (d - 1), seed=5001 + j).scale(10.0 ** (-3 * j))
        y = t if y is None else y + t
    eps = 1e-5
else:
    x = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2001); y = x + x
    eps = 1e-8
for _ in range(3):
    z = y.clone().round(eps)
torch.cuda.synchronize()
z = y.clone(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
t = time.time(); z.round(eps); torch.cuda.synchronize()
print("ms", 1e3 * (time.time() - t), z.ranks(), z.last_round)
torch.cuda.cudart().cudaProfilerStop()
