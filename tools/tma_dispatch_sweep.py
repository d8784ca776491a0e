"""Which inner-product sweep kernel is faster per shape?  (TTB_INNER_TMA=0: three-phase kernel, 2: TMA strip kernel)"""
import os, sys, torch
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain
def ms(a, b, reps=5):
    for _ in range(2): a.inner_dev(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): a.inner_dev(b)
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
for (r, n, d) in [(256, 32, 16), (224, 32, 16), (192, 32, 16), (160, 32, 16), (128, 32, 16), (96, 64, 16), (256, 24, 16), (256, 16, 16), (256, 8, 16), (256, 64, 12), (200, 48, 16), (64, 128, 16)]:
    a = TensorTrain.rand([n] * d, [r] * (d - 1), seed=1); b = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2)
    out = {}
    for mode in ("0", "2"):
        os.environ["TTB_INNER_TMA"] = mode
        out[mode] = ms(a, b)
    os.environ.pop("TTB_INNER_TMA")
    dflt = ms(a, b)
    print(f"r={r:4d} n={n:4d} d={d}: three-phase {out['0']:7.3f} ms  tma {out['2']:7.3f} ms  default {dflt:7.3f} ms  tiles {n * r // 56}")
