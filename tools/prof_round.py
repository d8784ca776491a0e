import sys, torch, time
sys.path.insert(0,'.')
from tensor_networks_b200 import TensorTrain
d,n,r=6,64,128
x=TensorTrain.rand([n]*d,[r]*(d-1),seed=2001); y=x+x
z=y.clone().round(1e-8); torch.cuda.synchronize()
z=y.clone(); torch.cuda.synchronize(); t=time.time(); z.round(1e-8); torch.cuda.synchronize(); print("ms",1e3*(time.time()-t), z.ranks(), z.last_round)
