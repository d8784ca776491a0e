"""Single-train inner products over bond ranks and mode sizes (device resident): which path wins where.
Run with TTB_INNER_FUSED=0 (per-GEMM path only), TTB_INNER_TMA=2 (strip kernel wherever possible) or default."""
import sys, torch
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain
shapes = [(12, n, r) for r in (192, 208, 224, 240, 256) for n in (16, 24, 32, 48, 64)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split(',')) for s in sys.argv[1:]]
out = []
for (d, n, r) in shapes:
    a = TensorTrain.rand([n] * d, [r] * (d - 1), seed=1); b = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2)
    for _ in range(3): v = a.inner_dev(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): v = a.inner_dev(b)
    e1.record(); e1.synchronize()
    out.append("%d,%d,%d:%.3f" % (d, n, r, e0.elapsed_time(e1) / 10))
    del a, b
print(" ".join(out))
