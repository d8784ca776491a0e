import sys, torch, time
sys.path.insert(0,'.')
from tensor_networks_b200 import TensorTrain
x = TensorTrain.rand([16]*6, [16,64,64,64,16], seed=3001)
dense = x.dense_dev()
t=time.time()
tt = TensorTrain.from_dense(dense, 1e-10)
torch.cuda.synchronize(); print("ranks", tt.ranks(), time.time()-t)
