"""Single-train inner products over bond ranks (device resident): which path wins where (TTB_INNER_FUSED / TTB_INNER_TMA)."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain
for (d, n, r) in [(20, 20, 240), (20, 20, 256), (20, 8, 256), (12, 64, 384), (12, 128, 384), (8, 64, 512), (8, 32, 640), (8, 64, 640), (6, 32, 1024), (20, 16, 256), (10, 12, 248)]:
    a = TensorTrain.rand([n] * d, [r] * (d - 1), seed=1); b = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2)
    for _ in range(3): v = a.inner_dev(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): v = a.inner_dev(b)
    e1.record(); e1.synchronize()
    print(d, n, r, "ms %.3f" % (e0.elapsed_time(e1) / 10))
