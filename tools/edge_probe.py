import sys, numpy as np, copy
sys.path.insert(0,'.')
from oracle import tt_oracle as orc
from tensor_networks_b200 import TensorTrain
rng=np.random.default_rng(1)
# zero tensor
x=orc.rand_tt([4,5,6],[3,3],rng); x[1]=x[1]*0.0
try:
    ref,_=orc.svd_round(copy.deepcopy(x),1e-8); print("oracle zero ranks", orc.ranks_of(ref))
except Exception as e: print("oracle zero raised", type(e).__name__, e)
try:
    t=TensorTrain.from_cores(copy.deepcopy(x)).round(1e-8); print("dev zero ranks", t.ranks(), float(t.norm()))
except Exception as e: print("dev zero raised", type(e).__name__, e)
# modes of size 1
x=orc.rand_tt([1,7,1,5],[1,3,3],rng); y=orc.tt_add(x,x)
ref,_=orc.svd_round(copy.deepcopy(y),1e-10); t=TensorTrain.from_cores(copy.deepcopy(y)).round(1e-10)
print("n=1 modes", orc.ranks_of(ref), t.ranks(), np.abs(t.dense()-orc.to_dense(y)).max())
t2=TensorTrain.from_cores(copy.deepcopy(y)).gramsvd_round(1e-6); r2,_=orc.gramsvd_round(copy.deepcopy(y),1e-6)
print("gramsvd n=1 modes", orc.ranks_of(r2), t2.ranks())
