import sys, torch, time
sys.path.insert(0, '.')
from tensor_networks_b200.batch import TensorTrainBatch
B, d, n, r = 8192, 20, 8, 32
a = TensorTrainBatch.rand(B, [n] * d, [r] * (d - 1), seed=1)
b = TensorTrainBatch.rand(B, [n] * d, [r] * (d - 1), seed=2)
for _ in range(3):
    v = a.inner(b)
torch.cuda.synchronize()
t = time.time(); v = a.inner(b); torch.cuda.synchronize(); print("ms", 1e3 * (time.time() - t))
