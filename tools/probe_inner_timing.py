import os, sys, torch
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain
d, n, r = int(os.environ.get("PROBE_D", "64")), 32, 256
a = TensorTrain.rand([n] * d, [r] * (d - 1), seed=1); b = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2)
for _ in range(3):
    v = a.inner(b)
torch.cuda.synchronize()
print("inner", float(v))
