"""cfg5-shaped batched rounding: time per launch; TTB_BROUND_TIMING=1 prints the in-kernel phase clocks of CTA 0."""
import sys, torch
sys.path.insert(0, '.')
from tensor_networks_b200.batch import TensorTrainBatch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d, n, r = 20, 8, 16
x = TensorTrainBatch.rand(B, [n] * d, [r] * (d - 1), seed=1)
y = x + x
for _ in range(2):
    z = y.clone().round(1e-8)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    z = y.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); z.round(1e-8); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
print("B", B, "round ms", min(ts), "ranks", z.item_ranks[0].tolist()[:4])
