"""Steady-state TT-SVD of the cfg4 tensor for ncu (--profile-from-start off)."""
import sys, torch, time
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain
n, d, ranks = 16, 7, (16, 64, 64, 64, 64, 16)
x = TensorTrain.rand([n] * d, list(ranks), seed=3001)
dense = x.dense_dev(); del x
for _ in range(2):
    tt = TensorTrain.from_dense(dense, 1e-10)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
t = time.time(); tt = TensorTrain.from_dense(dense, 1e-10); torch.cuda.synchronize()
print("ms", 1e3 * (time.time() - t), tt.ranks())
torch.cuda.cudart().cudaProfilerStop()
