import sys, torch, time
sys.path.insert(0,'.')
from tensor_networks_b200.batch import TensorTrainBatch
B,d,n,r=592,20,8,16
x=TensorTrainBatch.rand(B,[n]*d,[r]*(d-1),seed=1); y=x+x
z=y.clone().round(1e-8); torch.cuda.synchronize()
z=y.clone(); torch.cuda.synchronize(); t=time.time(); z.round(1e-8); torch.cuda.synchronize(); print("ms",1e3*(time.time()-t), "per item-wave us", 1e3*(time.time()-t)/4*1e0)
