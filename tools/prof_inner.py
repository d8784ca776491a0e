"""cfg2 TT inner product (d=64, n=32, r=256) resident in HBM: CUDA-event time per sweep."""
import sys, torch
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain
d, n, r = 64, 32, 256
a = TensorTrain.rand([n] * d, [r] * (d - 1), seed=1)
b = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2)
for _ in range(3):
    v = a.inner_dev(b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    v = a.inner_dev(b)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 10
print("inner ms", ms, "TFLOP/s", 133.152407552 / ms, "value", float(v))
