import sys, torch, time
sys.path.insert(0, '.')
from tensor_networks_b200.batch import TensorTrainBatch
d, n, r = 20, 8, 32
for B in (8192, 4096, 2048, 1024, 600):
    a = TensorTrainBatch.rand(B, [n] * d, [r] * (d - 1), seed=1)
    b = TensorTrainBatch.rand(B, [n] * d, [r] * (d - 1), seed=2)
    for _ in range(3):
        v = a.inner(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): v = a.inner(b)
    e1.record(); e1.synchronize()
    print("B", B, "ms", e0.elapsed_time(e1) / 5)
    del a, b
